"""Generate golden input/output vectors by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the reference is not shipped to the GPU box):
    python tests/golden/make_golden.py
Writes small .npz fixtures next to this file; tests/test_oracle_golden.py checks the oracle against them
(and, when /root/reference is present, re-derives them live).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.ref_import import import_reference  # noqa: E402


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def gen_pixel_ops(ref):
    g = torch.Generator().manual_seed(0)
    out = {}
    N, H, W, C = 2, 12, 10, 12
    # --- interpolate_bilinear incl. out-of-range, exact-edge and integer coordinates
    data = torch.randn(7, 9, 3, generator=g)
    sx = torch.rand(40, generator=g) * 10 - 1
    sy = torch.rand(40, generator=g) * 8 - 1
    sx[:6] = torch.tensor([0.0, 8.0, 8.0, 3.0, -1e-3, 8.001])
    sy[:6] = torch.tensor([0.0, 6.0, 2.5, 6.0, 2.0, 2.0])
    out.update(ib_data=data, ib_sx=sx, ib_sy=sy, ib_out=ref.misc.interpolate_bilinear(data, sx.clone(), sy.clone()))
    # --- TextureMapper (4 levels, apply_sh)
    tm = ref.network.TextureMapper(texture_size=16, texture_num_ch=C, mipmap_level=3, apply_sh=True)
    for i, t in enumerate(tm.textures):
        t.data = torch.randn(t.shape, generator=g) * 0.5
        out['tm_tex%d' % i] = t.data.clone()
    uv = torch.rand(N, H, W, 2, generator=g)
    uv[0, 0, 0] = torch.tensor([0.0, 0.0]); uv[0, 0, 1] = torch.tensor([1.0, 1.0]); uv[0, 0, 2] = torch.tensor([1.0, 0.0])
    sh = torch.randn(N, H, W, 9, generator=g)
    out.update(tm_uv=uv, tm_sh=sh, tm_out=tm(uv.clone(), sh, sh_start_ch=3), tm_flat=tm.flatten_mipmap(0, 6))
    # --- RaySampler, both modes
    rs = ref.network.RaySampler(num_azi=6, num_polar=2, interval_polar=5)
    rsd = ref.network.RaySampler(num_azi=6, num_polar=2, interval_polar=10, mode='diffuse')
    TBN = torch.linalg.qr(torch.randn(N, H, W, 3, 3, generator=g))[0]
    vdt = torch.nn.functional.normalize(torch.randn(N, H, W, 3, generator=g), dim=-1)
    alpha = (torch.rand(N, H, W, 1, generator=g) > 0.3).float()
    TBN = TBN * alpha[..., None]
    d0, u0, t0 = rs(TBN, vdt, alpha)
    d1, u1, _ = rsd(TBN, vdt, alpha)
    out.update(rs_Rs=rs.Rs, rs_piv=rs.pivots_dir, rsd_Rs=rsd.Rs, rsd_piv=rsd.pivots_dir, rs_TBN=TBN, rs_vdt=vdt, rs_alpha=alpha,
               rs_dir=d0, rs_uv=u0, rs_tan=t0, rsd_dir=d1, rsd_uv=u1)
    # --- RayRenderer (seperate albedo, 13 + 13 rays)
    rays_uv = torch.cat((u0, u1), -1)
    R = rays_uv.shape[-1]
    rays_lt = torch.rand(N, R, 3, H, W, generator=g) * 2
    lp = torch.rand(1, 8, 16, 3, generator=g) * 3
    alb_s = torch.rand(N, 3, H, W, generator=g)
    alb_d = torch.rand(N, 3, H, W, generator=g)
    rr = ref.network.RayRenderer(None, ref.network.Interpolater())
    o = rr(alb_s, rays_uv, rays_lt, lp=lp, albedo_diffuse=alb_d, num_ray_diffuse=13, seperate_albedo=True)
    out.update(rr_uv=rays_uv, rr_lt=rays_lt, rr_lp=lp, rr_alb_s=alb_s, rr_alb_d=alb_d, rr_out=o[0], rr_out_s=o[1], rr_out_d=o[2],
               rr_ltt_s=o[3], rr_ltt_d=o[4], rr_color=o[5])
    o2 = rr(alb_s, rays_uv, rays_lt, lp=lp, num_ray_diffuse=13, seperate_albedo=False)
    out.update(rr2_out=o2[0])
    # --- RaysLTChromLoss
    img = torch.rand(N, 3, H, W, generator=g) * 0.1
    l, chrom, mean, diff = ref.network.RaysLTChromLoss()(rays_lt, alpha.permute(0, 3, 1, 2), img)
    out.update(cl_img=img, cl_loss=l, cl_chrom=chrom, cl_mean=mean, cl_diff=diff)
    # --- spherical mapping + inverse, reconstruct / fit
    dirs = torch.nn.functional.normalize(torch.randn(3, 50, generator=g), dim=0)
    out.update(sm_dirs=dirs, sm_uv=ref.render.spherical_mapping(dirs))
    uvg = torch.rand(2, 30, generator=g)
    uvg[0, :3] = torch.tensor([0.0, 1.0, 0.5])
    out.update(smi_uv=uvg, smi_dir=ref.render.spherical_mapping_inv(uvg))
    basis = torch.randn(50, 9, generator=g)
    coeff = torch.randn(2, 9, 3, generator=g)
    out.update(sh_basis=basis, sh_coeff=coeff, sh_recon=ref.sph_harm.reconstruct_sh(coeff, basis),
               sh_recon2=ref.sph_harm.reconstruct_sh(coeff[0], basis),
               sh_fit=ref.sph_harm.fit_sh_coeff(torch.randn(2, 50, 3, generator=torch.Generator().manual_seed(5)), basis))
    out.update(sh_fit_samples=torch.randn(2, 50, 3, generator=torch.Generator().manual_seed(5)))
    # --- view dir map, reflect dir, TBN map
    K = torch.tensor([[[30.0, 0, 5.0], [0, 30.0, 6.0], [0, 0, 1]]]).repeat(N, 1, 1)
    Rm = torch.linalg.qr(torch.randn(N, 3, 3, generator=g))[0]
    vd, vdc = ref.camera.get_view_dir_map((H, W), torch.inverse(K), Rm.transpose(1, 2))
    out.update(vd_Kinv=torch.inverse(K), vd_Rinv=Rm.transpose(1, 2), vd_out=vd, vd_cam=vdc)
    nf = 20
    faces_v = torch.randn(nf, 3, 3, generator=g)
    faces_vt = torch.rand(nf, 3, 2, generator=g)
    fim = torch.randint(-1, nf, (N, H, W), generator=g).int()
    nrm = torch.nn.functional.normalize(torch.randn(N, H, W, 3, generator=g), dim=-1)
    out.update(tbn_faces_v=faces_v, tbn_faces_vt=faces_vt, tbn_fim=fim, tbn_normal=nrm,
               tbn_out=ref.render.get_TBN_map(nrm, fim, faces_v=faces_v, faces_texcoord=faces_vt))
    return out


def gen_unet(ref):
    torch.manual_seed(0)
    net = ref.network.RenderingNet(nf0=4, in_channels=5, out_channels=3, num_down_unet=5, out_channels_gcn=8, use_gcn=True)
    net.eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.train()
    x = torch.randn(1, 5, 32, 32)
    with torch.no_grad():
        y = net(x, torch.zeros(1, 8))
    sd = {k: v for k, v in net.state_dict().items() if 'fuse' not in k}
    out = {'x': x, 'y': y, 'keys': np.array(sorted(net.state_dict().keys()))}
    for k, v in sd.items():
        out['sd/' + k] = v
    return out


def main():
    ref = import_reference()
    np.savez_compressed(os.path.join(HERE, 'pixel_ops.npz'), **_np(gen_pixel_ops(ref)))
    np.savez_compressed(os.path.join(HERE, 'unet_small.npz'), **_np(gen_unet(ref)))
    for f in ('pixel_ops.npz', 'unet_small.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
