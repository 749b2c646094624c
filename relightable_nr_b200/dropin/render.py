"""Drop-in for the reference's ``render`` module (render.py:11-220)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from . import misc, sph_harm

vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
_lib.register_sigs({
    "rnr_face_tangents": [vp, vp, vp, vp, i32, vp],
    "rnr_tbn_map": [vp, vp, vp, vp, vp, i64, i32, vp],
    "rnr_interp_vertex_attr": [vp, i32, i32, i32, vp, i32, vp, vp, vp, i32, i64, vp],
})


def _s():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t, name):
    if not t.is_cuda:
        raise TypeError('%s must be a CUDA tensor (librnr_b200 has no CPU path)' % name)


def interp_vertex_attr(v_attr, faces_v_idx, face_index_map, weight_map):
    """v_attr [nv,A] or [N,nv,A]; faces_v_idx [N,nf,3]; face_index_map [N,H,W]; weight_map [N,H,W,3,1] -> [N,H,W,A] (render.py:11-28)."""
    _need_cuda(v_attr, 'v_attr')
    if v_attr.dim() == 2:
        v_attr = v_attr[None]
    N, H, W = face_index_map.shape
    attr = v_attr.float().contiguous()
    faces = faces_v_idx.to(torch.int32).contiguous()
    if faces.shape[0] != N:
        faces = faces.expand(N, -1, -1).contiguous()
    fidx = face_index_map.to(torch.int32).contiguous()
    w = weight_map.float().reshape(N, H, W, 3).contiguous()
    A = attr.shape[-1]
    out = torch.empty((N, H, W, A), dtype=torch.float32, device=attr.device)
    _lib.check(_lib.lib().rnr_interp_vertex_attr(attr.data_ptr(), attr.shape[0], attr.shape[1], A, faces.data_ptr(), faces.shape[1],
                                                 fidx.data_ptr(), w.data_ptr(), out.data_ptr(), N, H * W, _s()), 'rnr_interp_vertex_attr')
    return out


def texture_mapping(texture, uv_map):
    """texture [H,W,C], uv_map [N,H,W,2] -> [N,H,W,C] (render.py:31-46; like the reference it rescales ``uv_map`` in place)."""
    th, tw = float(texture.shape[0]), float(texture.shape[1])
    uv_map[..., 0] = uv_map[..., 0] * (tw - 1)
    uv_map[..., 1] = th - 1 - uv_map[..., 1] * (th - 1)
    return misc.interpolate_bilinear(texture, uv_map[..., 0], uv_map[..., 1])


def lp_mapping(lp, dir_map, alpha_map):
    """Sample an equirect light probe [H,W,C] along directions [3,...] (render.py:49-59; the reference's body refers to an
    undefined name ``render`` and cannot run -- this is what it evidently intends)."""
    uv = spherical_mapping(dir_map)
    uv = uv * alpha_map - (alpha_map == 0).to(dir_map.dtype)
    return misc.interpolate_bilinear(lp, uv[0] * float(lp.shape[1] - 1), uv[1] * float(lp.shape[0] - 1))


def sample_light_dir(azi_deg, pol_deg):
    """Grid of light directions: returns (world-space [3,S] with y up / z out, z-up [3,S]) (render.py:62-84)."""
    azi, pol = torch.meshgrid([azi_deg, pol_deg], indexing='ij')
    azi, ele = azi * np.pi / 180.0, np.pi / 2.0 - pol * np.pi / 180.0
    x, y, z = sph_harm.sph2cart(azi, ele, 1.0)
    zup = torch.nn.functional.normalize(torch.stack((x, y, z), 0), dim=0)
    world = torch.stack((zup[0], zup[2], -zup[1]), 0)
    return world.flatten(1), zup.flatten(1)


def _equirect_uv(x, y, z, dim):
    return torch.stack((torch.atan2(z, x) * 0.5 / np.pi + 0.5, torch.acos(y) * 1.0 / np.pi), dim=dim)


def spherical_mapping(l_dir):
    """[3,...] -> equirect uv [2,...]: u = atan2(z,x)/2pi + .5, v = acos(y)/pi (render.py:87-93).  Set-up-time helper; the
    per-pixel version is fused into the ray-sampler kernel."""
    return _equirect_uv(l_dir[0], l_dir[1], l_dir[2], 0)


def spherical_mapping_batch(l_dir):
    """[N,3,...] -> [N,2,...] (render.py:96-102)."""
    return _equirect_uv(l_dir[:, 0], l_dir[:, 1], l_dir[:, 2], 1)


def spherical_mapping_inv(lp_samples_uv):
    """equirect uv [2,S] -> unit direction [3,S] (render.py:105-121), with exact zeros of z at u = 0 and u = 1."""
    y = torch.cos(lp_samples_uv[1] * np.pi)
    rxz = (1 - y ** 2).sqrt()
    t = lp_samples_uv[0] * 2 - 1
    x = rxz * torch.cos(t * np.pi)
    z = rxz * torch.sin(t * np.pi)
    keep = ((t != 1.0).to(rxz.dtype) * 2 - 1) * ((t != -1.0).to(rxz.dtype) * 2 - 1)
    return torch.nn.functional.normalize(torch.stack((x, y, z * keep), 0), dim=0)


def get_TBN_map(normal_map, face_index_map, faces_v=None, faces_texcoord=None, tangent=None):
    """normal_map [N,H,W,3], face_index_map [N,H,W], faces_v [nf,3,3], faces_texcoord [nf,3,2] (or per-face ``tangent`` [nf,3])
    -> TBN [N,H,W,3,3] with columns (tangent, bitangent, normal) (render.py:124-168).  Raises ValueError('nan value detected')
    like the reference -- checked with one flag read-back instead of three isnan().sum() syncs."""
    _need_cuda(normal_map, 'normal_map')
    dev = normal_map.device
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    if tangent is None:
        assert faces_v is not None and faces_texcoord is not None
        fv = faces_v.to(dev).float().contiguous()
        fvt = faces_texcoord.to(dev).float().contiguous()
        nf = fv.shape[0]
        tan = torch.empty((nf, 3), dtype=torch.float32, device=dev)
        _lib.check(L.rnr_face_tangents(fv.data_ptr(), fvt.data_ptr(), tan.data_ptr(), flag.data_ptr(), nf, _s()), 'rnr_face_tangents')
    else:
        tan = torch.nn.functional.normalize(tangent.to(dev).float(), dim=-1).contiguous()
        nf = tan.shape[0]
    N, H, W = face_index_map.shape
    nm = normal_map.float().contiguous()
    fidx = face_index_map.to(dev).to(torch.int32).contiguous()
    tbn = torch.empty((N, H, W, 3, 3), dtype=torch.float32, device=dev)
    _lib.check(L.rnr_tbn_map(nm.data_ptr(), fidx.data_ptr(), tan.data_ptr(), tbn.data_ptr(), flag.data_ptr(), N * H * W, nf, _s()),
               'rnr_tbn_map')
    if int(flag.item()) != 0:
        raise ValueError('nan value detected')
    return tbn


def get_TBN_map_perpixel(normal_map, position_map, uv_map, alpha_map):
    """Screen-space TBN from finite differences of position / uv (render.py:171-220).  No call sites in the reference scripts;
    kept for API completeness as tensor arithmetic."""
    a = alpha_map
    data = torch.cat((position_map, uv_map), -1)
    zx, zy = torch.zeros_like(a[:, :, :1]), torch.zeros_like(a[:, :1])
    ax0 = ((torch.cat((a[:, :, 1:], zx), 2) * a) != 0).to(normal_map.dtype)
    ax1 = ((ax0 == 0) & (a != 0)).to(normal_map.dtype)
    ay0 = ((torch.cat((a[:, 1:], zy), 1) * a) != 0).to(normal_map.dtype)
    ay1 = ((ay0 == 0) & (a != 0)).to(normal_map.dtype)
    ex = data[:, :, 1:] - data[:, :, :-1]
    padx = torch.zeros_like(data[:, :, :1])
    ex = ax0 * torch.cat((ex, padx), 2) + ax1 * torch.cat((padx, ex), 2)
    ey = data[:, 1:] - data[:, :-1]
    pady = torch.zeros_like(data[:, :1])
    ey = ay0 * torch.cat((ey, pady), 1) + ay1 * torch.cat((pady, ey), 1)
    dp1, duv1, dp2, duv2 = ex[..., :3], ex[..., 3:], ey[..., :3], ey[..., 3:]
    f = 1.0 / (duv1[..., 0] * duv2[..., 1] - duv2[..., 0] * duv1[..., 1])
    t = torch.nn.functional.normalize(f[..., None] * (duv2[..., 1:2] * dp1 - duv1[..., 1:2] * dp2), dim=-1)
    b = torch.nn.functional.normalize(f[..., None] * (-duv2[..., 0:1] * dp1 + duv1[..., 0:1] * dp2), dim=-1)
    return torch.stack((t, b, normal_map), dim=4)
