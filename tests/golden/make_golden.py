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


def _load_ref_file(name, relpath):
    """Import one pure-Python file of the reference's neural_renderer package without its CUDA-extension imports."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join('/root/reference/neural_renderer/neural_renderer', relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_raster():
    """projection / vertices_to_faces from the reference's Python; the z-buffer from the reference's own CUDA kernel bodies
    compiled for the CPU (oracle/_ref/libref_raster.so, see oracle/build_oracle.py)."""
    from oracle import build_oracle, raster as Rr, ref_raster
    build_oracle.build(verbose=False)
    proj_mod = _load_ref_file('ref_nr_projection', 'projection.py')
    v2f_mod = _load_ref_file('ref_nr_v2f', 'vertices_to_faces.py')
    g = torch.Generator().manual_seed(7)
    out = {}
    # --- projection with distortion, offset and scale, batch 2
    N, nv = 2, 50
    verts = torch.randn(1, nv, 3, generator=g) * 0.5
    K = torch.tensor([[[80.0, 0.3, 31.0], [0, 82.0, 33.0], [0, 0, 1]]]).repeat(N, 1, 1)
    Rm = torch.linalg.qr(torch.randn(N, 3, 3, generator=g))[0]
    t = torch.tensor([[[0.1, -0.2, 3.0]], [[-0.1, 0.05, 2.5]]])
    dist = torch.tensor([[0.05, -0.01, 0.002, -0.003, 0.001], [0.0, 0.0, 0.0, 0.0, 0.0]])
    off = torch.tensor([[1.0, -2.0], [0.0, 0.0]])
    sc = torch.tensor([[0.9, 1.1], [1.0, 1.0]])
    out.update(pj_v=verts, pj_K=K, pj_R=Rm, pj_t=t, pj_dist=dist, pj_off=off, pj_sc=sc,
               pj_out=proj_mod.projection(verts.repeat(N, 1, 1), K, Rm, t, dist, 64, off, sc),
               pj_out_plain=proj_mod.projection(verts.repeat(N, 1, 1), K, Rm, t, torch.zeros(N, 5), 64, None, None))
    faces = torch.randint(0, nv, (1, 30, 3), generator=g).int()
    out.update(vf_faces=faces, vf_out=v2f_mod.vertices_to_faces(out['pj_out'], faces),
               vf_attr=v2f_mod.vertex_attrs_to_faces(torch.randn(1, nv, 2, generator=torch.Generator().manual_seed(9)), faces))
    out['vf_attr_in'] = torch.randn(1, nv, 2, generator=torch.Generator().manual_seed(9))
    # --- z-buffer of a UV sphere seen by a spiral camera + a soup of random triangles (ties, slivers, back faces), 48 px
    import math
    m = Rr.uv_sphere(12, 24)
    size = 48
    azi, ele = math.radians(-20.0), math.radians(15.0)
    pos = np.array([3 * math.cos(ele) * math.sin(azi), 3 * math.sin(ele), 3 * math.cos(ele) * math.cos(azi)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, [0, 1, 0]); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    Rc = np.stack([right, down, fwd]).astype(np.float32)
    pose = torch.eye(4)[None].clone()
    pose[0, :3, :3] = torch.from_numpy(Rc)
    pose[0, :3, 3] = torch.from_numpy(-Rc @ pos.astype(np.float32))
    Kc = torch.tensor([[[1.2 * size, 0, size / 2], [0, 1.2 * size, size / 2], [0, 0, 1]]])
    uvz = proj_mod.projection(torch.from_numpy(m['v'])[None], Kc, pose[:, :3, :3], pose[:, :3, -1, None].permute(0, 2, 1),
                              torch.zeros(1, 5), size, None, None)
    f_sphere = v2f_mod.vertices_to_faces(uvz, torch.from_numpy(m['f'])[None]).numpy()
    soup = (torch.rand(1, 60, 3, 3, generator=g) * 2 - 1)
    soup[..., 2] = soup[..., 2] * 0.5 + 1.5
    soup[0, 10] = soup[0, 9]                        # exact duplicate: depth tie -> lowest index wins
    soup[0, 20, :, 2] = -1.0                        # behind the near plane
    soup[0, 21, :, 2] = 2e5                         # beyond far
    f_soup = soup.numpy()
    for nm, f in (('sphere', f_sphere), ('soup', f_soup)):
        fim, wm, dm, fiv, finv = ref_raster.forward_face_index_map(f, size, 0.0, 1e5)
        out.update({'zb_%s_faces' % nm: f, 'zb_%s_fim' % nm: fim, 'zb_%s_wm' % nm: wm, 'zb_%s_dm' % nm: dm,
                    'zb_%s_finv' % nm: np.nan_to_num(finv, nan=0.0, posinf=0.0, neginf=0.0)})
    out.update(zb_size=np.array(size), zb_pose=pose, zb_K=Kc, zb_mesh_v=m['v'], zb_mesh_vt=m['vt'], zb_mesh_vn=m['vn'], zb_mesh_f=m['f'])
    return out


def gen_geometry(ref):
    """render.get_TBN_map (render.py:124-168), camera.get_view_dir_map (camera.py:5-32) and network.LightingLP (network.py:631-699,
    construction: cv2.resize INTER_AREA + bilinear sampling of the probes at the light directions) run by the real reference."""
    g = torch.Generator().manual_seed(5)
    out = {}
    N, H, W, nf = 2, 9, 11, 40
    faces_v = torch.randn(nf, 3, 3, generator=g)
    faces_vt = torch.rand(nf, 3, 2, generator=g)
    # orient every face's texture triangle so that the reference's clamp(min=1e-8) on the uv determinant is not the decisive term
    det = (faces_vt[:, 1, 0] - faces_vt[:, 0, 0]) * (faces_vt[:, 2, 1] - faces_vt[:, 0, 1]) - \
          (faces_vt[:, 2, 0] - faces_vt[:, 0, 0]) * (faces_vt[:, 1, 1] - faces_vt[:, 0, 1])
    flip = det < 0
    faces_vt[flip] = faces_vt[flip][:, [0, 2, 1]]
    faces_v[flip] = faces_v[flip][:, [0, 2, 1]]
    normal = torch.randn(N, H, W, 3, generator=g)
    normal[0, 0, :3] = 0.0                                    # uncovered pixels: zero normal -> zero TBN
    fidx = torch.randint(0, nf, (N, H, W), generator=g).int()
    out.update(tbn_faces_v=faces_v, tbn_faces_vt=faces_vt, tbn_normal=normal, tbn_fidx=fidx,
               tbn_out=ref.render.get_TBN_map(normal.clone(), fidx, faces_v, faces_vt))
    K = torch.tensor([[[14.0, 0, 5.5], [0, 13.0, 4.5], [0, 0, 1]], [[20.0, 0.3, 6.0], [0, 21.0, 4.0], [0, 0, 1]]])
    R = torch.linalg.qr(torch.randn(2, 3, 3, generator=g))[0]
    vd, vdc = ref.camera.get_view_dir_map((H, W), torch.inverse(K), R.transpose(1, 2).contiguous())
    out.update(vd_Kinv=torch.inverse(K), vd_Rinv=R.transpose(1, 2).contiguous(), vd_out=vd, vd_cam=vdc)
    l_dir = torch.nn.functional.normalize(torch.randn(3, 64, generator=g), dim=0)
    probes = [{'lp_img': torch.rand(1, 3, 30, 60, generator=g) * 4}, {'lp_img': torch.rand(1, 3, 25, 50, generator=g)}]
    lp = ref.network.LightingLP(l_dir, lp_dataloader=probes, lp_img_h=20, lp_img_w=40)
    out.update(lp_l_dir=l_dir, lp_probe0=probes[0]['lp_img'], lp_probe1=probes[1]['lp_img'], lp_l_samples=lp.l_samples.data,
               lp_lps=lp.lps, lp_uv=lp.l_samples_uv)
    return out


def gen_gcn(ref):
    """network.DenseDeepGCN (network.py:256-315) at a small size, training mode (batch-statistic BN, spectral-norm power
    iteration), stochastic dilation off so that the run is deterministic.  The state dict is captured BEFORE the forward."""
    import types
    torch.manual_seed(0)
    opt = types.SimpleNamespace(n_filters=16, kernel_size=4, act_type='relu', norm_type='batch', bias=True, epsilon=0.0,
                                stochastic=False, conv_type='edge', n_blocks=4, num_v_gcn=96, out_channels_gcn=8, in_channels=6,
                                block_type='res')
    net = ref.network.DenseDeepGCN(opt)
    net.train()
    g = torch.Generator().manual_seed(3)
    v = torch.randn(96, 3, generator=g)
    sd = {k: t.detach().clone() for k, t in net.state_dict().items()}
    inputs = types.SimpleNamespace(pos=v, x=v)
    fea = net(inputs)
    out = {'gcn_sd__' + k: t for k, t in sd.items()}
    out.update(gcn_v=v, gcn_fea=fea.detach(), gcn_keys=np.array(sorted(sd.keys())))
    return out


def gen_sphere_samples():
    """The reference's only in-tree data fixture: the 4096 quadrature directions every script loads
    (train_rnr.py:167, test_rnr.py, precompute.py) -- stored as float32 [4096,3]."""
    import scipy.io
    from tests.golden.ref_import import REF
    return {'sphere_samples': scipy.io.loadmat(os.path.join(REF, 'sphere_samples_4096.mat'))['sphere_samples'].astype(np.float32)}


def main():
    ref = import_reference()
    np.savez_compressed(os.path.join(HERE, 'sphere_samples_4096.npz'), **gen_sphere_samples())
    np.savez_compressed(os.path.join(HERE, 'gcn_small.npz'), **_np(gen_gcn(ref)))
    np.savez_compressed(os.path.join(HERE, 'geometry.npz'), **_np(gen_geometry(ref)))
    np.savez_compressed(os.path.join(HERE, 'pixel_ops.npz'), **_np(gen_pixel_ops(ref)))
    np.savez_compressed(os.path.join(HERE, 'unet_small.npz'), **_np(gen_unet(ref)))
    np.savez_compressed(os.path.join(HERE, 'raster.npz'), **_np(gen_raster()))
    for f in ('pixel_ops.npz', 'unet_small.npz', 'raster.npz', 'gcn_small.npz', 'geometry.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
