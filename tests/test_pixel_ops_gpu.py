"""GPU parity of the per-pixel CUDA operators against (a) golden vectors produced by the real reference
(tests/golden/pixel_ops.npz) and (b) the oracle's autograd for gradients.  Tolerances: max-abs 1e-5 for
sampled values / directions, 1e-4 relative for colours and gradients (atomic accumulation order)."""
import os

import numpy as np
import pytest
import torch

from oracle import pixel_ops as P

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def g():
    z = np.load(os.path.join(G, 'pixel_ops.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def close(a, b, tol=1e-5):
    a = a.detach().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=tol, atol=tol), (a - b).abs().max().item()


def test_interpolate_bilinear(g):
    from relightable_nr_b200 import ops
    out = ops.interpolate_bilinear(g['ib_data'].cuda(), g['ib_sx'].cuda(), g['ib_sy'].cuda())
    close(out, g['ib_out'])
    # gradient w.r.t. data
    d = g['ib_data'].clone().requires_grad_(True)
    P.interpolate_bilinear(d, g['ib_sx'], g['ib_sy']).square().sum().backward()
    dc = g['ib_data'].cuda().requires_grad_(True)
    ops.interpolate_bilinear(dc, g['ib_sx'].cuda(), g['ib_sy'].cuda()).square().sum().backward()
    close(dc.grad, d.grad, 1e-4)


def test_texture_mapper_fwd_bwd(g):
    from relightable_nr_b200 import ops
    tex = [g['tm_tex%d' % i] for i in range(3)]
    texc = [t.cuda().requires_grad_(True) for t in tex]
    out = ops.texture_mapper(texc, g['tm_uv'].cuda(), g['tm_sh'].cuda(), sh_start_ch=3)
    close(out, g['tm_out'])
    Rw = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * Rw.cuda()).sum().backward()
    texo = [t.clone().requires_grad_(True) for t in tex]
    (P.texture_mapper_forward(texo, g['tm_uv'], g['tm_sh'], 3) * Rw).sum().backward()
    for a, b in zip(texc, texo):
        close(a.grad, b.grad, 1e-4)
    # untouched texels keep an exactly-zero gradient (albedo-mean loss relies on it, train_rnr.py:598)
    assert ((texo[0].grad == 0) == (texc[0].grad.cpu() == 0)).all()
    flat = ops.flatten_mipmap(texc, 0, 6)
    close(flat, g['tm_flat'])
    for t in texc:
        t.grad = None
    Rf = torch.randn(flat.shape, generator=torch.Generator().manual_seed(4))
    (flat * Rf.cuda()).sum().backward()
    for t in texo:
        t.grad = None
    (P.flatten_mipmap(texo, 0, 6) * Rf).sum().backward()
    for a, b in zip(texc, texo):
        close(a.grad, b.grad, 1e-4)


def test_ray_sampler(g):
    from relightable_nr_b200 import ops
    d, uv, tan = ops.ray_sampler(g['rs_piv'].cuda(), g['rs_TBN'].cuda(), g['rs_vdt'].cuda(), g['rs_alpha'].cuda(), True)
    close(d, g['rs_dir']); close(uv, g['rs_uv']); close(tan, g['rs_tan'])
    d, uv, _ = ops.ray_sampler(g['rsd_piv'].cuda(), g['rs_TBN'].cuda(), None, g['rs_alpha'].cuda(), False)
    close(d, g['rsd_dir']); close(uv, g['rsd_uv'])


def test_ray_renderer_fwd_bwd(g):
    from relightable_nr_b200 import ops
    ins = {k: g[k].cuda().requires_grad_(True) for k in ('rr_alb_s', 'rr_alb_d', 'rr_lt', 'rr_lp')}
    o = ops.ray_render(ins['rr_alb_s'], g['rr_uv'].cuda(), ins['rr_lt'], ins['rr_lp'], albedo_diffuse=ins['rr_alb_d'],
                       num_ray_diffuse=13, seperate_albedo=True)
    for a, k in zip(o, ('rr_out', 'rr_out_s', 'rr_out_d', 'rr_ltt_s', 'rr_ltt_d', 'rr_color')):
        close(a, g[k], 1e-5)
    Rw = torch.randn(o[0].shape, generator=torch.Generator().manual_seed(5))
    (o[0] * Rw.cuda()).sum().backward()
    ino = {k: g[k].clone().requires_grad_(True) for k in ins}
    oo = P.ray_renderer_forward(ino['rr_alb_s'], g['rr_uv'], ino['rr_lt'], ino['rr_lp'], albedo_diffuse=ino['rr_alb_d'],
                                num_ray_diffuse=13, seperate_albedo=True)
    (oo[0] * Rw).sum().backward()
    for k in ins:
        close(ins[k].grad, ino[k].grad, 1e-4)
    o2 = ops.ray_render(g['rr_alb_s'].cuda(), g['rr_uv'].cuda(), g['rr_lt'].cuda(), g['rr_lp'].cuda(), num_ray_diffuse=13)
    close(o2[0], g['rr2_out'], 1e-5)


def test_chrom_loss_fwd_bwd(g):
    from relightable_nr_b200 import ops
    alpha = g['rs_alpha'].permute(0, 3, 1, 2).contiguous()
    lt = g['rr_lt'].cuda().requires_grad_(True)
    l, chrom, mean, diff = ops.chrom_loss(lt, alpha.cuda(), g['cl_img'].cuda())
    close(l, g['cl_loss']); close(chrom, g['cl_chrom']); close(mean, g['cl_mean']); close(diff, g['cl_diff'])
    (l * 3.0).backward()
    lo = g['rr_lt'].clone().requires_grad_(True)
    (P.rays_lt_chrom_loss(lo, alpha, g['cl_img'])[0] * 3.0).backward()
    close(lt.grad, lo.grad, 1e-4)


def test_sh_ops(g):
    from relightable_nr_b200 import ops
    close(ops.sh_reconstruct(g['sh_coeff'].cuda(), g['sh_basis'].cuda()), g['sh_recon'], 1e-5)
    close(ops.sh_reconstruct(g['sh_coeff'][0].cuda(), g['sh_basis'].cuda()), g['sh_recon2'], 1e-5)
    close(ops.sh_fit(g['sh_fit_samples'].cuda(), g['sh_basis'].cuda()), g['sh_fit'], 1e-5)
    c = g['sh_coeff'].cuda().requires_grad_(True)
    ops.sh_reconstruct(c, g['sh_basis'].cuda()).square().sum().backward()
    co = g['sh_coeff'].clone().requires_grad_(True)
    P.reconstruct_sh(co, g['sh_basis']).square().sum().backward()
    close(c.grad, co.grad, 1e-4)
    # degree-2 basis against the fp64 oracle (closed-form known answers are checked on CPU)
    d = torch.nn.functional.normalize(torch.randn(1000, 3, generator=torch.Generator().manual_seed(0)), dim=-1)
    Y = ops.sh_basis_l2(d.cuda())
    ref = torch.from_numpy(P.evaluate_sh_basis(2, d.numpy())).float()
    close(Y, ref, 1e-6)


def test_large_lmax10_reconstruct():
    """LightingSH.reconstruct_lp shape of train_rnr.py:271 (256x512 envmap, 121 coefficients)."""
    from relightable_nr_b200 import ops
    gen = torch.Generator().manual_seed(0)
    basis = torch.randn(256 * 512, 121, generator=gen)
    coeff = torch.randn(121, 3, generator=gen)
    out = ops.sh_reconstruct(coeff.cuda(), basis.cuda())
    ref = basis.double() @ coeff.double()
    assert (out.cpu().double() - ref).abs().max() < 1e-4
