"""The RNR / DNR per-view step assembled from the drop-in modules, exactly as the reference scripts do it.

``RNRPipeline.train_step`` is the body of train_rnr.py:490-623 (forward, the four losses, backward, Adam) and
``RNRPipeline.render`` the body of test_rnr.py:335-371; ``DNRPipeline`` is train_dnr.py:240-275.  They exist so that
bench.py, the smoke test and the parity tests drive the same call sequence a user's unchanged script would, without
needing the (absent) material_sphere dataset on disk.  ``synthetic_view`` builds the per-view maps of a unit sphere
analytically -- the geometry the rasterizer produces for the material_sphere proxy -- for a spiral camera.
"""
import math

import numpy as np
import torch

from .dropin import camera as _camera
from .dropin import network as _network
from .dropin import sph_harm as _sph_harm


def fibonacci_sphere(n, device='cpu'):
    """[3, n] near-uniform unit directions (stand-in for the reference's sphere_samples_4096.mat quadrature nodes)."""
    i = torch.arange(n, dtype=torch.float64) + 0.5
    phi = math.pi * (1 + 5 ** 0.5) * i
    z = 1 - 2 * i / n
    r = torch.sqrt(1 - z * z)
    return torch.stack((r * torch.cos(phi), r * torch.sin(phi), z)).float().to(device)


def _view_dir_map(H, W, K, R, device):
    """camera.get_view_dir_map (camera.py:5-32) in plain torch ops (any device): the ray through each pixel centre,
    -K^-1 [u+.5, v+.5, 1] normalised, rotated to world space (R: world->camera rotation) and normalised -- points to the camera."""
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device) + 0.5, torch.arange(W, dtype=torch.float32, device=device) + 0.5,
                          indexing='ij')
    pix = torch.stack((u, v, torch.ones_like(u)), -1)                        # [H,W,3]
    cam = torch.nn.functional.normalize(-(pix @ torch.inverse(K).t()), dim=-1)
    return torch.nn.functional.normalize(cam @ R, dim=-1)                    # R^T applied to row vectors


def _sh_basis_l2(d):
    """sph_harm.evaluate_sh_basis(lmax=2) in closed form (real, orthonormal, no Condon-Shortley phase) -> [...,9]."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    return torch.stack((torch.full_like(x, 0.28209479177387814), 0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x,
                        1.0925484305920792 * x * y, 1.0925484305920792 * y * z, 0.31539156525252005 * (3 * z * z - 1),
                        1.0925484305920792 * x * z, 0.5462742152960396 * (x * x - y * y)), -1)


def synthetic_view(img_size=512, view_idx=0, device='cuda', radius=3.0, focal=None, seed=0):
    """Per-view maps of a unit sphere at the origin seen from spiral camera ``view_idx``: the dict a
    ``ViewDataset`` item holds after precompute.py (dataio.py:219-245) -- uv_map, sh_basis_map, normal_map, view_dir_map,
    view_dir_map_tangent, TBN_map, alpha_map, img_gt -- as fp32 tensors with batch dimension 1 on ``device``.  Plain torch
    arithmetic on any device: bench.py feeds the SAME tensors to the GPU arm and to the CPU reference arm."""
    H = W = int(img_size)
    focal = focal if focal is not None else 1.2 * img_size
    azi = math.radians(-2.0 * view_idx)
    ele = math.radians(0.125 * view_idx)
    pos = np.array([radius * math.cos(ele) * math.sin(azi), radius * math.sin(ele), radius * math.cos(ele) * math.cos(azi)])
    RT = _camera.RT_from_pos_lookat(pos)
    R = torch.tensor(RT[:3, :3], dtype=torch.float32, device=device)
    K = torch.tensor([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]], dtype=torch.float32, device=device)
    view_dir = _view_dir_map(H, W, K, R, device)[None]                       # [1,H,W,3], points to the camera
    o = torch.tensor(pos, dtype=torch.float32, device=device)
    d = -view_dir[0]
    b = (d * o).sum(-1)
    disc = b * b - ((o * o).sum() - 1.0)
    hit = disc > 0
    t = -b - torch.sqrt(disc.clamp(min=0))
    p = torch.nn.functional.normalize(o + t[..., None] * d, dim=-1)
    alpha = hit.float()
    normal = p * alpha[..., None]
    u = torch.atan2(p[..., 2], p[..., 0]) / (2 * math.pi) + 0.5
    v = torch.acos(p[..., 1].clamp(-1, 1)) / math.pi
    uv = torch.stack((u, 1 - v), -1) * alpha[..., None]
    # tangent along +u (longitude), bitangent = n x t, re-orthogonalised like render.get_TBN_map
    tan = torch.stack((-p[..., 2], torch.zeros_like(u), p[..., 0]), -1)
    tan = torch.nn.functional.normalize(tan + 1e-6 * torch.tensor([1.0, 0, 0], device=device), dim=-1)
    bit = torch.nn.functional.normalize(torch.cross(p, tan, dim=-1), dim=-1)
    tan = torch.nn.functional.normalize(torch.cross(bit, p, dim=-1), dim=-1)
    TBN = torch.stack((tan, bit, p), dim=-1) * alpha[..., None, None]
    vdt = torch.einsum('hwji,hwj->hwi', TBN, view_dir[0])
    sh = _sh_basis_l2(view_dir[0])
    g = torch.Generator(device='cpu').manual_seed(seed + view_idx)
    img = torch.rand((1, 3, H, W), generator=g).to(device) * alpha[None, None]
    return {
        'uv_map': uv[None].contiguous(), 'sh_basis_map': sh[None].contiguous(), 'normal_map': normal[None].contiguous(),
        'view_dir_map': view_dir.contiguous(), 'view_dir_map_tangent': vdt[None].contiguous(), 'TBN_map': TBN[None].contiguous(),
        'alpha_map': alpha[None].contiguous(), 'img_gt': img.contiguous(),
    }


class RNRPipeline:
    """Module set of train_rnr.py:245-376 with its defaults (texture 512^2 x 24 ch x 4 mips, nf0 64, lmax 10, 13+13 rays,
    256x512 SH envmap) and the step of train_rnr.py:490-623."""

    def __init__(self, device='cuda', img_size=512, texture_size=512, texture_num_ch=24, mipmap_level=4, nf0=64, sh_lmax=10,
                 num_l_samples=4096, lp_recon_h=256, lp_recon_w=512, lr=1e-3, seed=0, loss_weights=None, dropout=True,
                 capturable=False, l_dir=None):
        self.device = torch.device(device)
        self.img_size = img_size
        torch.manual_seed(seed)
        # light directions [3, S]: the caller's (the reference loads sphere_samples_4096.mat, train_rnr.py:167-169) or a Fibonacci sphere
        l_dir = fibonacci_sphere(num_l_samples) if l_dir is None else torch.as_tensor(l_dir, dtype=torch.float32)
        num_l_samples = l_dir.shape[1]
        self.l_dir = l_dir.clone()
        self.interpolater = _network.Interpolater()
        self.texture_mapper = _network.TextureMapper(texture_size, texture_num_ch, mipmap_level, apply_sh=True)
        g = torch.Generator().manual_seed(seed + 11)
        init = torch.randn((1, (sh_lmax + 1) ** 2, 3), generator=g) * 0.1
        init[:, 0] = 1.0
        self.lighting_model = _network.LightingSH(l_dir, lmax=sh_lmax, num_lighting=1, num_channel=3, init_coeff=init,
                                                  lp_recon_h=lp_recon_h, lp_recon_w=lp_recon_w)
        self.ray_sampler = _network.RaySampler(num_azi=6, num_polar=2, interval_polar=5)
        self.ray_sampler_diffuse = _network.RaySampler(num_azi=6, num_polar=2, interval_polar=10, mode='diffuse')
        self.num_ray_total = self.ray_sampler.num_ray + self.ray_sampler_diffuse.num_ray
        self.render_net = _network.RenderingNet(nf0=nf0, in_channels=self.num_ray_total * 3 + 6 + texture_num_ch,
                                                out_channels=3 * self.num_ray_total, num_down_unet=5)
        self.render_net.set_input_grad_channels(self.num_ray_total * 3 + 6, self.num_ray_total * 3 + 6 + texture_num_ch)
        self.ray_renderer = _network.RayRenderer(self.lighting_model, self.interpolater)
        self.chrom_loss = _network.RaysLTChromLoss()
        self.modules = [self.texture_mapper, self.lighting_model, self.ray_sampler, self.ray_sampler_diffuse, self.render_net,
                        self.ray_renderer]
        for m in self.modules:
            m.to(self.device)
            m.train()
        if not dropout:
            for m in self.render_net.modules():
                if isinstance(m, torch.nn.Dropout2d):
                    m.eval()
        with torch.no_grad():
            for lvl, t in enumerate(self.texture_mapper.textures):       # 0.5 randn texture around the init value
                t.add_(0.05 * torch.randn(t.shape, generator=torch.Generator().manual_seed(seed + 20 + lvl)).to(self.device))
        self.w = dict(lighting=1.0, lighting_uncovered=0.1, rays_lt_chrom=1.0, alb=1.0)
        if loss_weights:
            self.w.update(loss_weights)
        # lighting-loss targets (train_rnr.py:307-339): samples of the initial envmap, all covered
        with torch.no_grad():
            # (targets are samples of the stitched envmap, not of its SH fit: never bit-equal to the estimate, so |.|' is defined)
            self.l_samples_init = _sph_harm.reconstruct_sh(self.lighting_model.coeff.data[0], self.lighting_model.basis_val).clone()
            self.l_samples_init += 0.05 * torch.randn(self.l_samples_init.shape, generator=torch.Generator().manual_seed(seed + 31)).to(self.device)
            self.l_samples_init_mask = torch.ones(num_l_samples, dtype=torch.bool, device=self.device)
            self.l_samples_init_mask[::7] = False
        params = list(self.texture_mapper.parameters()) + list(self.lighting_model.parameters()) + list(self.render_net.parameters())
        # torch.optim.Adam like train_rnr.py:376; ``fused=True`` only picks torch's single-kernel multi-tensor implementation
        self.optimizer = torch.optim.Adam(params, lr=lr, fused=True, capturable=capturable)
        self.optimizer.zero_grad()
        self.lighting_idx = 0

    # ---- forward of train_rnr.py:512-547 / test_rnr.py:335-371 --------------------------------------------------------
    def forward(self, view):
        alpha_map = view['alpha_map'][:, None]                                   # [N,1,H,W]
        N, _, H, W = alpha_map.shape
        neural_img = self.texture_mapper(view['uv_map'], view['sh_basis_map'], sh_start_ch=6)
        albedo_diffuse, albedo_specular = neural_img[:, :3], neural_img[:, 3:6]
        alpha_last = alpha_map.permute(0, 2, 3, 1)
        rays_dir, rays_uv, _ = self.ray_sampler(view['TBN_map'], view['view_dir_map_tangent'], alpha_last)
        rays_dir_d, rays_uv_d, _ = self.ray_sampler_diffuse(view['TBN_map'], view['view_dir_map_tangent'], alpha_last)
        num_ray_diffuse = rays_uv_d.shape[-1]
        rays_dir = torch.cat((rays_dir, rays_dir_d), dim=-1)
        rays_uv = torch.cat((rays_uv, rays_uv_d), dim=-1)
        R = rays_uv.shape[-1]
        net_in = torch.cat((rays_dir.permute(0, -1, -2, 1, 2).reshape(N, -1, H, W), view['normal_map'].permute(0, 3, 1, 2),
                            view['view_dir_map'].permute(0, 3, 1, 2), neural_img), dim=1)
        rays_lt = self.render_net(net_in, None).reshape(N, R, -1, H, W)
        rays_lt = (rays_lt * 0.5 + 0.5) * 2.0
        out = self.ray_renderer(albedo_specular, rays_uv, rays_lt, lighting_idx=self.lighting_idx, albedo_diffuse=albedo_diffuse,
                                num_ray_diffuse=num_ray_diffuse, seperate_albedo=True)
        return out[0], rays_lt, alpha_map

    def losses(self, view, final, rays_lt, alpha_map):
        coeff = self.lighting_model.get_lighting_params(self.lighting_idx)
        l_est = _sph_harm.reconstruct_sh(coeff, self.lighting_model.basis_val)
        # train_rnr.py:578-579 (mask products instead of boolean indexing: same sums, no host sync)
        m = self.l_samples_init_mask.float()[:, None]
        d = (self.l_samples_init - l_est).abs()
        loss_lighting = (d * m).sum() / m.sum() * self.w['lighting'] + (d * (1 - m)).sum() / (1 - m).sum() * self.w['lighting_uncovered']
        a = alpha_map[:, :, 5:-5, 5:-5]
        img_gt = view['img_gt']
        loss_rn = torch.nn.functional.l1_loss((final[:, :, 5:-5, 5:-5] * a).contiguous().view(-1), (img_gt[:, :, 5:-5, 5:-5] * a).reshape(-1))
        loss_chrom = self.chrom_loss(rays_lt, alpha_map, img_gt)[0] * self.w['rays_lt_chrom']
        tm = self.texture_mapper
        loss_alb = 0
        for c0 in (3, 0):
            tex = tm.flatten_mipmap(start_ch=c0, end_ch=c0 + 3)
            valid = (tex != tm.tex_flatten_mipmap_init[..., c0:c0 + 3]).any(dim=-1, keepdim=True).to(tex.dtype)
            cnt = valid.sum(dim=(0, 1, 2))
            loss_alb = loss_alb + ((tex * valid).sum(dim=(0, 1, 2)) / cnt.clamp(min=1) - 0.5).abs().sum() / 3 * (cnt > 0).float()
        loss_alb = loss_alb * self.w['alb']
        return loss_lighting + loss_rn + loss_chrom + loss_alb, dict(lighting=loss_lighting, rn=loss_rn, chrom=loss_chrom, alb=loss_alb)

    @property
    def fused(self):
        """The same iteration with its per-pixel stages fused around the U-Net (relightable_nr_b200/fused.py)."""
        if getattr(self, '_fused', None) is None:
            from .fused import FusedRNRStep
            self._fused = FusedRNRStep(self)
        return self._fused

    def train_step(self, view, step_optimizer=True, fused=False):
        if fused:
            loss, final = self.fused.train_step(view, step_optimizer=step_optimizer)
            return loss.detach(), final
        final, rays_lt, alpha_map = self.forward(view)
        loss, parts = self.losses(view, final, rays_lt, alpha_map)
        loss.backward()
        if step_optimizer:
            self.optimizer.step()
            self.optimizer.zero_grad()
        return loss.detach(), final.detach()

    def make_graphed_step(self, example_view, grad_hook=None, warmup=3, fused=False):
        """Capture ``train_step`` (forward, four losses, backward, [grad_hook], Adam) into ONE CUDA graph.

        Returns ``(step, static_view)``: ``step(view)`` copies the per-view maps into the static input buffers, replays the
        graph and returns the (static) loss tensor.  ~400 kernel launches per iteration collapse into one graph launch, which
        removes the Python / launch overhead that otherwise bounds the step once the kernels are fast.  ``grad_hook(params)``,
        if given, runs between backward and the optimiser step inside the capture (the data-parallel all-reduce).  The
        pipeline must have been built with ``capturable=True``."""
        dev = self.device
        static_view = {k: v.detach().clone() for k, v in example_view.items()}
        params = [p for grp in self.optimizer.param_groups for p in grp['params']]

        def body_fused():
            # grad_hook receives the step's gradient buffers (engine flat buffer, texture levels, SH coefficients)
            if grad_hook is not None:
                self.fused.grad_hook = lambda gs: grad_hook(gs)
            loss, _ = self.fused.train_step(static_view)
            return loss.detach()

        def body():
            if fused:
                return body_fused()
            self.optimizer.zero_grad(set_to_none=True)
            final, rays_lt, alpha_map = self.forward(static_view)
            loss, _ = self.losses(static_view, final, rays_lt, alpha_map)
            loss.backward()
            if grad_hook is not None:
                grad_hook(params)
            self.optimizer.step()
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # builds every plan / workspace and warms the allocator outside the capture
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.lib().rnr_launch_count()
        with torch.cuda.graph(graph):
            static_loss = body()
        self._graph = graph
        self.graph_launches = int(_lib.lib().rnr_launch_count() - n0)     # librnr_b200 kernels recorded in the graph

        def step(view=None):
            if view is not None and view is not static_view:
                for k, v in view.items():
                    static_view[k].copy_(v, non_blocking=True)
            graph.replay()
            return static_loss

        return step, static_view

    @torch.no_grad()
    def render(self, view, fused=False):
        if fused:
            return self.fused.render(view)
        return self.forward(view)[0]


class DNRPipeline:
    """train_dnr.py:138-193, 240-275: TextureMapper(C ch, SH on channels 3..11) -> RenderingNet(use_gcn=False) -> masked L1."""

    def __init__(self, device='cuda', img_size=512, texture_size=512, texture_num_ch=16, mipmap_level=4, nf0=80, lr=1e-3, seed=0):
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.texture_mapper = _network.TextureMapper(texture_size, texture_num_ch, mipmap_level, apply_sh=True)
        self.render_net = _network.RenderingNet(nf0=nf0, in_channels=texture_num_ch, out_channels=3, num_down_unet=5, use_gcn=False)
        for m in (self.texture_mapper, self.render_net):
            m.to(self.device)
            m.train()
        params = list(self.texture_mapper.parameters()) + list(self.render_net.parameters())
        self.optimizer = torch.optim.Adam(params, lr=lr)

    def forward(self, view):
        neural_img = self.texture_mapper(view['uv_map'], view['sh_basis_map'])
        return (self.render_net(neural_img, None) * 0.5 + 0.5) * 2.0

    def train_step(self, view):
        out = self.forward(view)
        a = view['alpha_map'][:, None, 5:-5, 5:-5]
        loss = torch.nn.functional.l1_loss((out[:, :, 5:-5, 5:-5] * a).reshape(-1), (view['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.detach(), out.detach()
