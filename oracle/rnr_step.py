"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of one RNR / DNR iteration of the reference.

Follows /root/reference/train_rnr.py:512-608 (forward, the four losses; autograd gives the backward of :618) and
/root/reference/train_dnr.py:240-262, composed from the golden-pinned pieces in oracle/pixel_ops.py and oracle/unet.py.
Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product path.
"""
import torch
import torch.nn.functional as F

from . import pixel_ops as P
from .unet import rendering_net_forward


def rnr_forward(textures, unet_sd, coeff, basis_val_recon, lp_hw, pivots_s, pivots_d, view, drop=None, sh_start_ch=6):
    """train_rnr.py:512-539.  ``textures`` list of [1,S,S,C]; ``unet_sd`` state_dict of RenderingNet ('net.' prefix);
    ``coeff`` [121,3]; ``view`` dict of per-view maps (batch-first).  Returns (final [N,3,H,W], rays_lt, alpha [N,1,H,W])."""
    alpha_map = view['alpha_map'][:, None]
    N, _, H, W = alpha_map.shape
    neural_img = P.texture_mapper_forward(textures, view['uv_map'], view['sh_basis_map'], sh_start_ch)
    albedo_diffuse, albedo_specular = neural_img[:, :3], neural_img[:, 3:6]
    a_last = alpha_map.permute(0, 2, 3, 1)
    d0, uv0, _ = P.ray_sampler_forward(pivots_s, view['TBN_map'], view['view_dir_map_tangent'], a_last, 'reflect')
    d1, uv1, _ = P.ray_sampler_forward(pivots_d, view['TBN_map'], view['view_dir_map_tangent'], a_last, 'diffuse')
    rays_dir = torch.cat((d0, d1), -1)
    rays_uv = torch.cat((uv0, uv1), -1)
    R = rays_uv.shape[-1]
    net_in = torch.cat((rays_dir.permute(0, 4, 3, 1, 2).reshape(N, -1, H, W), view['normal_map'].permute(0, 3, 1, 2),
                        view['view_dir_map'].permute(0, 3, 1, 2), neural_img), 1)
    rays_lt = rendering_net_forward(unet_sd, net_in, drop=drop).reshape(N, R, -1, H, W)
    rays_lt = (rays_lt * 0.5 + 0.5) * 2.0
    lp = P.reconstruct_sh(coeff, basis_val_recon).reshape(lp_hw[0], lp_hw[1], -1)[None]
    out = P.ray_renderer_forward(albedo_specular, rays_uv, rays_lt, lp, albedo_diffuse=albedo_diffuse,
                                 num_ray_diffuse=d1.shape[-1], seperate_albedo=True)
    return out[0], rays_lt, alpha_map


def rnr_losses(textures, tex_init, coeff, basis_val, l_init, l_mask, view, final, rays_lt, alpha_map, w):
    """train_rnr.py:558-608 with the weights dict ``w`` (lighting, lighting_uncovered, rays_lt_chrom, alb)."""
    l_est = P.reconstruct_sh(coeff, basis_val)
    loss_lighting = (l_init[l_mask] - l_est[l_mask]).abs().sum() / l_mask.float().sum() * w['lighting'] + \
                    (l_init[~l_mask] - l_est[~l_mask]).abs().sum() / (~l_mask).float().sum() * w['lighting_uncovered']
    a = alpha_map[:, :, 5:-5, 5:-5]
    loss_rn = F.l1_loss((final[:, :, 5:-5, 5:-5] * a).reshape(-1), (view['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
    loss_chrom = P.rays_lt_chrom_loss(rays_lt, alpha_map, view['img_gt'])[0] * w['rays_lt_chrom']
    loss_alb = 0
    for c0 in (3, 0):
        tex = P.flatten_mipmap(textures, c0, c0 + 3)
        valid = (tex != tex_init[..., c0:c0 + 3]).any(-1, keepdim=True).to(tex.dtype)
        if valid.sum() > 0:
            loss_alb = loss_alb + ((tex * valid).sum((0, 1, 2)) / valid.sum((0, 1, 2)) - 0.5).abs().sum() / 3
    return loss_lighting + loss_rn + loss_chrom + loss_alb * w['alb']


def dnr_forward(textures, unet_sd, view, drop=None):
    """train_dnr.py:252-257: texture (SH on channels 3..11) -> U-Net -> tanh -> (x*0.5+0.5)*2."""
    neural_img = P.texture_mapper_forward(textures, view['uv_map'], view['sh_basis_map'], 3)
    return (rendering_net_forward(unet_sd, neural_img, drop=drop) * 0.5 + 0.5) * 2.0


def sh_tables(l_dir, lmax, lp_hw):
    """LightingSH's two basis tables (network.py:557, 577-581) from the ORACLE's evaluate_sh_basis (scipy-pinned,
    tests/test_oracle_golden.py): ``l_dir`` [3,S] float32 -> (basis_val [S,B], basis_val_recon [H*W,B]) float32."""
    import numpy as np
    basis_val = torch.from_numpy(P.evaluate_sh_basis(lmax, l_dir.t().contiguous().numpy())).float()
    h, w = lp_hw
    vv, uu = torch.meshgrid(torch.arange(h, dtype=torch.float32) / (h - 1), torch.arange(w, dtype=torch.float32) / (w - 1), indexing='ij')
    grid_dir = P.spherical_mapping_inv(torch.stack((uu, vv)).flatten(1)).t().contiguous()
    return basis_val, torch.from_numpy(P.evaluate_sh_basis(lmax, grid_dir.numpy())).float()


def state_from_pipeline(pipe):
    """CPU copies of everything the oracle needs from a relightable_nr_b200.pipeline.RNRPipeline (duck-typed: the
    oracle never imports the product).  Learnable state and inputs are copied; the SH basis tables are NOT taken from the
    pipeline under test -- they are rebuilt by the oracle from the light directions (``pipe.l_dir``)."""
    cpu = lambda t: t.detach().cpu().clone()
    lm = pipe.lighting_model
    lp_hw = (int(lm.lp_recon_h), int(lm.lp_recon_w))
    basis_val, basis_val_recon = sh_tables(cpu(pipe.l_dir).float(), int(lm.lmax), lp_hw)
    return dict(
        textures=[cpu(t) for t in pipe.texture_mapper.textures],
        tex_init=cpu(pipe.texture_mapper.tex_flatten_mipmap_init),
        unet_sd={k: cpu(v) for k, v in pipe.render_net.state_dict().items()},
        coeff=cpu(lm.coeff[pipe.lighting_idx]),
        basis_val=basis_val, basis_val_recon=basis_val_recon, lp_hw=lp_hw,
        pivots_s=cpu(pipe.ray_sampler.pivots_dir), pivots_d=cpu(pipe.ray_sampler_diffuse.pivots_dir),
        l_init=cpu(pipe.l_samples_init), l_mask=cpu(pipe.l_samples_init_mask), w=dict(pipe.w),
    )


def rnr_step(state, view, requires_grad=True, drop=None):
    """One oracle iteration on CPU: returns (loss, final, grads dict).  grads keys: 'textures.i', 'coeff', 'unet/<key>'."""
    view = {k: v.detach().cpu() for k, v in view.items()}
    tex = [t.clone().requires_grad_(requires_grad) for t in state['textures']]
    coeff = state['coeff'].clone().requires_grad_(requires_grad)
    sd = {k: (v.clone().requires_grad_(requires_grad) if (v.dtype.is_floating_point and 'running' not in k and v.dim() > 0) else v)
          for k, v in state['unet_sd'].items()}
    # Conv2dSame duplicates alias one parameter in the reference; the oracle reads the '.net.1.' names only
    final, rays_lt, alpha_map = rnr_forward(tex, sd, coeff, state['basis_val_recon'], state['lp_hw'], state['pivots_s'],
                                            state['pivots_d'], view, drop=drop)
    loss = rnr_losses(tex, state['tex_init'], coeff, state['basis_val'], state['l_init'], state['l_mask'], view, final, rays_lt,
                      alpha_map, state['w'])
    grads = {}
    if requires_grad:
        loss.backward()
        for i, t in enumerate(tex):
            grads['textures.%d' % i] = t.grad
        grads['coeff'] = coeff.grad
        for k, v in sd.items():
            if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None:
                grads['unet/' + k] = v.grad
    return loss.detach(), final.detach(), grads


def rnr_trajectory(state, views, lr=1e-3):
    """``len(views)`` consecutive iterations of train_rnr.py:490-623 on CPU in fp32: forward, the four losses, backward and
    ``torch.optim.Adam(lr)`` (train_rnr.py:376) over textures + SH coefficients + U-Net parameters, BatchNorm in batch-statistic
    mode, dropout off.  Returns (losses, final state tensors) -- the reference trajectory a GPU run must follow."""
    tex = [t.clone().requires_grad_(True) for t in state['textures']]
    coeff = state['coeff'].clone().requires_grad_(True)
    sd = {k: (v.clone().requires_grad_(True) if (v.dtype.is_floating_point and 'running' not in k and v.dim() > 0) else v)
          for k, v in state['unet_sd'].items()}
    leaves = tex + [coeff] + [v for v in sd.values() if isinstance(v, torch.Tensor) and v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=lr)
    losses = []
    for view in views:
        view = {k: v.detach().cpu() for k, v in view.items()}
        opt.zero_grad(set_to_none=True)
        final, rays_lt, alpha_map = rnr_forward(tex, sd, coeff, state['basis_val_recon'], state['lp_hw'], state['pivots_s'],
                                                state['pivots_d'], view)
        loss = rnr_losses(tex, state['tex_init'], coeff, state['basis_val'], state['l_init'], state['l_mask'], view, final, rays_lt,
                          alpha_map, state['w'])
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses, dict(textures=[t.detach() for t in tex], coeff=coeff.detach(),
                        unet_sd={k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()})
