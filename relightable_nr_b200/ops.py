"""Host-side operator wrappers (torch.autograd.Function + ctypes) over librnr_b200.so for the per-pixel part
of the hot path.  Each wrapper mirrors one reference operator (argument meaning, layouts, error behaviour)
so the drop-in modules in ``dropin/`` are thin.  CUDA tensors only; there is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)
_lib.register_sigs({
    "rnr_texmap_fwd": [_pp, _ip, i32, i32, vp, vp, i32, vp, i32, i32, i32, vp],
    "rnr_texmap_bwd": [_pp, _ip, i32, i32, vp, vp, i32, vp, i32, i32, i32, vp],
    "rnr_flatten_mipmap": [_pp, _pp, _ip, i32, i32, i32, i32, vp, vp, i32, vp],
    "rnr_bilinear_fwd": [vp, i32, i32, i32, i32, vp, vp, vp, i64, i32, vp],
    "rnr_bilinear_bwd": [vp, i32, i32, i32, i32, vp, vp, vp, i64, i32, vp],
    "rnr_ray_sampler_fwd": [vp, vp, vp, vp, i32, i32, vp, vp, vp, i64, vp],
    "rnr_ray_render_fwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "rnr_ray_render_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                           i32, i32, i32, vp],
    "rnr_sh_basis_l2": [vp, vp, i64, vp],
    "rnr_sh_reconstruct": [vp, vp, vp, i64, i32, i32, i32, vp],
    "rnr_sh_project": [vp, vp, vp, i64, i32, i32, i32, f32, vp],
    "rnr_chrom_loss_fwd": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "rnr_chrom_loss_bwd": [vp, vp, vp, i32, i32, i32, i32, vp, vp, f32, vp, i32, vp],
    "rnr_l1_masked": [vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp],
    "rnr_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp],
})


def _s():
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError('%s must be a CUDA tensor (librnr_b200 has no CPU path)' % name)
    if t.device.index != torch.cuda.current_device():
        # kernels are launched on the CURRENT device's stream: a tensor living elsewhere would be dereferenced on the wrong GPU
        raise RuntimeError('%s lives on %s but the current CUDA device is cuda:%d; wrap the call in torch.cuda.device(%s.device)'
                           % (name, t.device, torch.cuda.current_device(), name))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t):
    return t.data_ptr() if t is not None else None


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors])
    return C.cast(arr, _pp), arr


def _int_array(vals):
    arr = (C.c_int * len(vals))(*vals)
    return C.cast(arr, _ip), arr


# ----------------------------------------------------------------------------------------------------
# misc.interpolate_bilinear (misc.py:5-42) / network.Interpolater (network.py:322-337)
# ----------------------------------------------------------------------------------------------------
class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, sx, sy):
        # data [Nd,Hd,Wd,C]; sx, sy [N, ...]
        Nd, Hd, Wd, Cc = data.shape
        N = sx.shape[0]
        M = sx[0].numel()
        out = torch.empty((*sx.shape, Cc), dtype=torch.float32, device=data.device)
        _lib.check(_lib.lib().rnr_bilinear_fwd(_p(data), Nd, Hd, Wd, Cc, _p(sx), _p(sy), _p(out), M, N, _s()), 'rnr_bilinear_fwd')
        ctx.save_for_backward(sx, sy)
        ctx.dshape = data.shape
        return out

    @staticmethod
    def backward(ctx, g):
        sx, sy = ctx.saved_tensors
        Nd, Hd, Wd, Cc = ctx.dshape
        gd = torch.zeros(ctx.dshape, dtype=torch.float32, device=g.device)
        g = g.contiguous()
        _lib.check(_lib.lib().rnr_bilinear_bwd(_p(gd), Nd, Hd, Wd, Cc, _p(sx), _p(sy), _p(g), sx[0].numel(), sx.shape[0], _s()),
                   'rnr_bilinear_bwd')
        return gd, None, None


def interpolate_bilinear(data, sub_x, sub_y):
    """data [H,W,C]; sub_x, sub_y [...] -> [..., C]   (misc.interpolate_bilinear)."""
    data = _cuda_f32(data, 'data')
    sx = _cuda_f32(sub_x, 'sub_x').to(data.device)
    sy = _cuda_f32(sub_y, 'sub_y').to(data.device)
    if sx.shape != sy.shape:
        sx, sy = torch.broadcast_tensors(sx, sy)
        sx, sy = sx.contiguous(), sy.contiguous()
    shape = sx.shape
    out = _Bilinear.apply(data[None], sx.reshape(1, -1), sy.reshape(1, -1))
    return out.reshape(*shape, data.shape[-1])


def interpolate_bilinear_batched(data, sub_x, sub_y):
    """data [N,H,W,C] or [1,H,W,C]; sub_x, sub_y [N, ...] -> [N, ..., C]   (network.Interpolater.forward)."""
    if data.shape[0] != 1 and data.shape[0] != sub_x.shape[0]:
        raise ValueError('data.shape[0] should be 1 or batch size')
    data = _cuda_f32(data, 'data')
    sx, sy = _cuda_f32(sub_x, 'sub_x'), _cuda_f32(sub_y, 'sub_y')
    shape = sx.shape
    out = _Bilinear.apply(data, sx.reshape(shape[0], -1), sy.reshape(shape[0], -1))
    return out.reshape(*shape, data.shape[-1])


# ----------------------------------------------------------------------------------------------------
# network.TextureMapper.forward (network.py:67-91)
# ----------------------------------------------------------------------------------------------------
class _TexMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, sh, sh_start, *textures):
        N, H, W, _ = uv.shape
        Cc = textures[0].shape[-1]
        sizes = [int(t.shape[1]) for t in textures]
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=uv.device)
        tp, _k1 = _ptr_array(textures)
        sp, _k2 = _int_array(sizes)
        _lib.check(_lib.lib().rnr_texmap_fwd(tp, sp, len(textures), Cc, _p(uv), _p(sh), sh_start, _p(out), N, H, W, _s()),
                   'rnr_texmap_fwd')
        ctx.save_for_backward(uv, sh if sh is not None else torch.empty(0, device=uv.device))
        ctx.has_sh = sh is not None
        ctx.meta = (sizes, Cc, sh_start, [t.shape for t in textures], [t.requires_grad for t in textures])
        return out

    @staticmethod
    def backward(ctx, g):
        uv, sh = ctx.saved_tensors
        sizes, Cc, sh_start, shapes, req = ctx.meta
        N, H, W, _ = uv.shape
        grads = [torch.zeros(s, dtype=torch.float32, device=g.device) if r else None for s, r in zip(shapes, req)]
        gp, _k1 = _ptr_array(grads)
        sp, _k2 = _int_array(sizes)
        g = g.contiguous()
        _lib.check(_lib.lib().rnr_texmap_bwd(gp, sp, len(grads), Cc, _p(uv), _p(sh) if ctx.has_sh else None, sh_start, _p(g),
                                             N, H, W, _s()), 'rnr_texmap_bwd')
        return (None, None, None, *grads)


def texture_mapper(textures, uv_map, sh_basis_map=None, sh_start_ch=3, apply_sh=True):
    """textures: sequence of [1,S_i,S_i,C] fp32 CUDA tensors -> neural image [N,C,H,W]."""
    uv = _cuda_f32(uv_map, 'uv_map')
    sh = _cuda_f32(sh_basis_map, 'sh_basis_map') if (apply_sh and sh_basis_map is not None) else None
    tex = [t if (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()) else _cuda_f32(t, 'texture') for t in textures]
    Cc = tex[0].shape[-1]
    if sh is not None and not (0 <= sh_start_ch and sh_start_ch + 9 <= Cc):
        raise ValueError('sh_start_ch + 9 exceeds the number of texture channels')
    return _TexMap.apply(uv, sh, int(sh_start_ch), *tex)


class _FlattenMipmap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c0, c1, *textures):
        sizes = [int(t.shape[1]) for t in textures]
        Cc = textures[0].shape[-1]
        S0 = sizes[0]
        out = torch.empty((1, S0, S0, c1 - c0), dtype=torch.float32, device=textures[0].device)
        tp, _k1 = _ptr_array(textures)
        gp, _k3 = _ptr_array([None] * len(textures))
        sp, _k2 = _int_array(sizes)
        _lib.check(_lib.lib().rnr_flatten_mipmap(tp, gp, sp, len(textures), Cc, c0, c1 - c0, _p(out), None, 0, _s()),
                   'rnr_flatten_mipmap')
        ctx.meta = (sizes, Cc, c0, c1, [t.shape for t in textures], [t.requires_grad for t in textures])
        return out

    @staticmethod
    def backward(ctx, g):
        sizes, Cc, c0, c1, shapes, req = ctx.meta
        grads = [torch.zeros(s, dtype=torch.float32, device=g.device) if r else None for s, r in zip(shapes, req)]
        tp, _k1 = _ptr_array([None] * len(grads))
        gp, _k3 = _ptr_array(grads)
        sp, _k2 = _int_array(sizes)
        g = g.contiguous()
        _lib.check(_lib.lib().rnr_flatten_mipmap(tp, gp, sp, len(grads), Cc, c0, c1 - c0, None, _p(g), 1, _s()),
                   'rnr_flatten_mipmap(bwd)')
        return (None, None, *grads)


def flatten_mipmap(textures, start_ch, end_ch):
    return _FlattenMipmap.apply(int(start_ch), int(end_ch), *textures)


# ----------------------------------------------------------------------------------------------------
# network.RaySampler.forward (network.py:445-472)
# ----------------------------------------------------------------------------------------------------
def ray_sampler(pivots_dir, TBN, view_dir_tangent, alpha, reflect):
    TBN = _cuda_f32(TBN, 'TBN_matrices')
    alpha = _cuda_f32(alpha, 'alpha_map')
    piv = _cuda_f32(pivots_dir, 'pivots_dir')
    R = piv.shape[1]
    lead = TBN.shape[:-2]
    P = 1
    for d in lead:
        P *= d
    rays_dir = torch.empty((*lead, 3, R), dtype=torch.float32, device=TBN.device)
    rays_uv = torch.empty((*lead, 2, R), dtype=torch.float32, device=TBN.device)
    if reflect:
        vdt = _cuda_f32(view_dir_tangent, 'view_dir_map_tangent')
        tan = torch.empty((*lead, 3, R), dtype=torch.float32, device=TBN.device)
    else:
        vdt, tan = None, None
    _lib.check(_lib.lib().rnr_ray_sampler_fwd(_p(TBN), _p(vdt), _p(alpha), _p(piv), R, 1 if reflect else 0, _p(rays_dir), _p(rays_uv),
                                              _p(tan), P, _s()), 'rnr_ray_sampler_fwd')
    return rays_dir, rays_uv, (tan if reflect else pivots_dir)


# ----------------------------------------------------------------------------------------------------
# network.RayRenderer.forward (network.py:481-527)
# ----------------------------------------------------------------------------------------------------
class _RayRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alb_s, alb_d, rays_uv, rays_lt, lp, Rd, no_albedo, separate):
        N, R = rays_lt.shape[0], rays_lt.shape[1]
        H, W = rays_lt.shape[3], rays_lt.shape[4]
        dev = rays_lt.device
        outs = [torch.empty((N, 3, H, W), dtype=torch.float32, device=dev) for _ in range(5)]
        col = torch.empty((N, R, 3, H, W), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().rnr_ray_render_fwd(_p(alb_s), _p(alb_d), _p(rays_uv), _p(rays_lt), _p(lp), lp.shape[0], lp.shape[1],
                                                 lp.shape[2], R, Rd, int(no_albedo), int(separate), *[_p(o) for o in outs], _p(col),
                                                 N, H, W, _s()), 'rnr_ray_render_fwd')
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(alb_s, alb_d if alb_d is not None else torch.empty(0, device=dev), rays_uv, rays_lt, lp, outs[3], outs[4])
        ctx.meta = (Rd, no_albedo, separate, alb_d is not None)
        ctx.mark_non_differentiable(col)
        return (*outs, col)

    @staticmethod
    def backward(ctx, g_out, g_os, g_od, g_ls, g_ld, g_col):
        alb_s, alb_d, rays_uv, rays_lt, lp, ltt_s, ltt_d = ctx.saved_tensors
        Rd, no_albedo, separate, has_d = ctx.meta
        N, R = rays_lt.shape[0], rays_lt.shape[1]
        H, W = rays_lt.shape[3], rays_lt.shape[4]
        dev = rays_lt.device
        need = ctx.needs_input_grad
        g_alb_s = torch.empty_like(alb_s) if need[0] else None
        g_alb_d = torch.empty_like(alb_d) if (has_d and need[1]) else None
        g_lt = torch.empty_like(rays_lt) if need[3] else None
        g_lp = torch.zeros_like(lp) if need[4] else None
        cg = [None if g is None else g.contiguous() for g in (g_out, g_os, g_od, g_ls, g_ld)]
        _lib.check(_lib.lib().rnr_ray_render_bwd(_p(alb_s), _p(alb_d) if has_d else None, _p(rays_uv), _p(rays_lt), _p(lp), lp.shape[0],
                                                 lp.shape[1], lp.shape[2], R, Rd, int(no_albedo), int(separate),
                                                 *[_p(g) for g in cg], _p(ltt_s), _p(ltt_d), _p(g_alb_s), _p(g_alb_d), _p(g_lt),
                                                 _p(g_lp), N, H, W, _s()), 'rnr_ray_render_bwd')
        return g_alb_s, g_alb_d, None, g_lt, g_lp, None, None, None


def ray_render(albedo_specular, rays_uv, rays_lt, lp, albedo_diffuse=None, num_ray_diffuse=0, no_albedo=False,
               seperate_albedo=False):
    """Returns (out, out_specular, out_diffuse, ltt_specular_map, ltt_diffuse_map, rays_color)."""
    alb_s = _cuda_f32(albedo_specular, 'albedo_specular')
    alb_d = _cuda_f32(albedo_diffuse, 'albedo_diffuse') if albedo_diffuse is not None else None
    return _RayRender.apply(alb_s, alb_d, _cuda_f32(rays_uv, 'rays_uv'), _cuda_f32(rays_lt, 'rays_lt'), _cuda_f32(lp, 'lp'),
                            int(num_ray_diffuse), bool(no_albedo), bool(seperate_albedo))


# ----------------------------------------------------------------------------------------------------
# sph_harm (sph_harm.py:41-102)
# ----------------------------------------------------------------------------------------------------
def sh_basis_l2(directions):
    """directions [..., 3] CUDA fp32 -> [..., 9] (degree-2 real orthonormal SH, reference ordering)."""
    d = _cuda_f32(directions, 'directions')
    out = torch.empty((*d.shape[:-1], 9), dtype=torch.float32, device=d.device)
    _lib.check(_lib.lib().rnr_sh_basis_l2(_p(d), _p(out), d.numel() // 3, _s()), 'rnr_sh_basis_l2')
    return out


class _SHReconstruct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff, basis):
        # coeff [L,B,C], basis [P,B]
        Lc, B, Cc = coeff.shape
        P = basis.shape[0]
        out = torch.empty((Lc, P, Cc), dtype=torch.float32, device=coeff.device)
        _lib.check(_lib.lib().rnr_sh_reconstruct(_p(basis), _p(coeff), _p(out), P, B, Cc, Lc, _s()), 'rnr_sh_reconstruct')
        ctx.save_for_backward(basis)
        ctx.cshape = coeff.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (basis,) = ctx.saved_tensors
        Lc, B, Cc = ctx.cshape
        gc = torch.zeros(ctx.cshape, dtype=torch.float32, device=g.device)
        g = g.contiguous()
        _lib.check(_lib.lib().rnr_sh_project(_p(basis), _p(g), _p(gc), basis.shape[0], B, Cc, Lc, 1.0, _s()), 'rnr_sh_project')
        return gc, None


def sh_reconstruct(sh_coeff, sh_basis_val):
    """[B,C] or [L,B,C] x [P,B] -> [P,C] or [L,P,C]   (sph_harm.reconstruct_sh)."""
    coeff = _cuda_f32(sh_coeff, 'sh_coeff')
    basis = _cuda_f32(sh_basis_val, 'sh_basis_val')
    if coeff.dim() == 2:
        return _SHReconstruct.apply(coeff[None], basis)[0]
    return _SHReconstruct.apply(coeff, basis)


def sh_fit(samples, sh_basis_val):
    """[P,C] or [L,P,C] x [P,B] -> [B,C] or [L,B,C]   (sph_harm.fit_sh_coeff, 4*pi/N quadrature)."""
    import math
    s = _cuda_f32(samples, 'samples')
    basis = _cuda_f32(sh_basis_val, 'sh_basis_val')
    two_d = s.dim() == 2
    if two_d:
        s = s[None]
    Lc, P, Cc = s.shape
    B = basis.shape[1]
    res = torch.zeros((Lc, B, Cc), dtype=torch.float32, device=s.device)
    _lib.check(_lib.lib().rnr_sh_project(_p(basis), _p(s.contiguous()), _p(res), P, B, Cc, Lc, 4.0 * math.pi / P, _s()), 'rnr_sh_project')
    return res[0] if two_d else res


# ----------------------------------------------------------------------------------------------------
# network.RaysLTChromLoss (network.py:395-411)
# ----------------------------------------------------------------------------------------------------
class _ChromLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_lt, alpha, img):
        N, R, _, H, W = rays_lt.shape
        dev = rays_lt.device
        chrom = torch.empty_like(rays_lt)
        mean = torch.empty((N, 1, 3, H, W), dtype=torch.float32, device=dev)
        diff = torch.empty((N, R, H, W), dtype=torch.float32, device=dev)
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().rnr_chrom_loss_fwd(_p(rays_lt), _p(alpha), _p(img), R, N, H, W, _p(chrom), _p(mean), _p(diff), _p(sums),
                                                 _s()), 'rnr_chrom_loss_fwd')
        loss = (sums[0] / sums[1] / R).float()
        ctx.save_for_backward(rays_lt, alpha, img if img is not None else torch.empty(0, device=dev), sums)
        ctx.has_img = img is not None
        ctx.mark_non_differentiable(chrom, mean, diff)
        return loss, chrom, mean, diff

    @staticmethod
    def backward(ctx, g_loss, g1, g2, g3):
        rays_lt, alpha, img, sums = ctx.saved_tensors
        N, R, _, H, W = rays_lt.shape
        g_lt = torch.empty_like(rays_lt)
        gl = g_loss.contiguous().float()
        _lib.check(_lib.lib().rnr_chrom_loss_bwd(_p(rays_lt), _p(alpha), _p(img) if ctx.has_img else None, R, N, H, W, _p(sums), _p(gl),
                                                 1.0, _p(g_lt), 0, _s()), 'rnr_chrom_loss_bwd')
        return g_lt, None, None


def chrom_loss(rays_lt, alpha_map, img=None):
    return _ChromLoss.apply(_cuda_f32(rays_lt, 'rays_lt'), _cuda_f32(alpha_map, 'alpha_map'),
                            _cuda_f32(img, 'img') if img is not None else None)
