"""The reference's UNCHANGED scripts, end to end, on the drop-in modules (north_star: "drops into train_rnr.py and test_rnr.py
unchanged"; SURVEY.md 8d "running the unchanged scripts end to end on the synthetic scene is a functional gate").

Scripts come from the git-ignored copy staged by tools/stage_reference.py (baseline/_ref/relightable-nr: the reference's own
files, byte for byte; it travels to the GPU box) and run through ``python -m relightable_nr_b200.run`` from the script
directory, exactly as a user of the reference would run them:

    precompute.py (mesh.obj, then mesh_7500v.obj --only_mesh_related)  ->  per-view maps on disk (dataio.py:219-245)
    train_rnr.py  --max_iter 4 --ckp_freq 2   (validation pass at iteration 0, checkpoints)      ->  model_*.pth + params.txt
    test_rnr.py   on test_seq/spiral_step720  ->  PNG renders
    train_dnr.py / test_dnr.py                ->  the same for the DNR baseline (nf0 = 80)

Parity of the rendered PNGs: the REAL reference modules (network.TextureMapper / RaySampler / RenderingNet / RayRenderer /
LightingLP / LightingSH, render.get_TBN_map, camera.get_view_dir_map -- imported from the staged copy and run on the CPU in
fp32) load the checkpoint the drop-in training wrote with ``strict=True`` and render the same views from G-buffers produced
by the oracle rasterizer.  Gate: PSNR >= 50 dB over the pixels whose coverage agrees (<= 0.05 % may differ: silhouette
pixels decided by an ulp of the projection), PNG 8-bit quantisation included (its own ceiling is 58.9 dB).
"""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.golden import ref_import
from tests.util import psnr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = ref_import.REF
SIZE = 512


def _run(script, *args, timeout=1500):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', 'relightable_nr_b200.run', os.path.join(REF, script)] + [str(a) for a in args],
                       cwd=REF, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout[-3000:] + '\n--- stderr ---\n' + r.stderr[-5000:])
    assert r.returncode == 0, '%s failed (rc %d):\n%s' % (script, r.returncode, tail)
    return r.stdout


@pytest.fixture(scope='module')
def scene(tmp_path_factory):
    if not ref_import.available():
        pytest.skip('reference scripts not staged (run tools/stage_reference.py in the build container)')
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_scene
    root = str(tmp_path_factory.mktemp('material_sphere'))
    info = make_scene.make_scene(root, n_views=4, n_test_views=2, img_size=SIZE, mesh_lat=128, mesh_lon=256)
    out = _run('precompute.py', '--data_root', root, '--gpu_id', '0', '--img_size', SIZE)
    assert 'View 0' in out
    _run('precompute.py', '--data_root', root, '--gpu_id', '0', '--img_size', SIZE, '--obj_fp', '_/mesh_7500v.obj', '--only_mesh_related', '1')
    return info


def _read_obj(path):
    """Minimal parser of the 'f v/vt/vn' OBJ make_scene writes -> the buffer dict of network.Rasterizer (network.py:129-134)."""
    v, vt, vn, f = [], [], [], []
    for ln in open(path):
        p = ln.split()
        if not p:
            continue
        if p[0] == 'v':
            v.append([float(x) for x in p[1:4]])
        elif p[0] == 'vt':
            vt.append([float(x) for x in p[1:3]])
        elif p[0] == 'vn':
            vn.append([float(x) for x in p[1:4]])
        elif p[0] == 'f':
            f.append([[int(i) - 1 for i in c.split('/')] for c in p[1:4]])
    f = np.asarray(f, dtype=np.int64)
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt)[None]
    return dict(vertices=t(v), vertices_texcoords=t(vt), vertices_normals=t(vn), faces=t(f[:, :, 0], torch.int32),
                faces_vt_idx=t(f[:, :, 1], torch.int32), faces_vn_idx=t(f[:, :, 2], torch.int32))


def _views(calib_fp, size):
    import scipy.io
    c = scipy.io.loadmat(calib_fp)
    out = []
    for i in range(c['poses'].shape[0]):
        pose = torch.from_numpy(c['poses'][i].astype(np.float32))[None]
        proj = torch.from_numpy(c['projs'][i].astype(np.float32))[None]
        out.append(dict(pose=pose, proj=proj, proj_inv=torch.inverse(proj), R_inv=pose[:, :3, :3].transpose(1, 2).contiguous()))
    return out


def _gbuffer(ref, mesh, vw, size):
    """Per-view maps of test_rnr.py:283-316 on the CPU: oracle rasterizer + the real render / camera / sph_harm functions."""
    from oracle.raster import rasterizer_forward
    r = rasterizer_forward(mesh, size, vw['proj'], vw['pose'])
    uv_map, alpha_map, fim, normal_map, faces_v, faces_vt = r[0], r[1], r[2], r[5], r[7], r[8]
    TBN = ref.render.get_TBN_map(normal_map, fim, faces_v=faces_v[0], faces_texcoord=faces_vt[0], tangent=None)
    view_dir, _ = ref.camera.get_view_dir_map(uv_map.shape[1:3], vw['proj_inv'], vw['R_inv'])
    vdt = torch.matmul(TBN.reshape(-1, 3, 3).transpose(-2, -1), view_dir.reshape(-1, 3, 1))[..., 0].reshape(view_dir.shape)
    vdt = torch.nn.functional.normalize(vdt, dim=-1)
    sh = torch.from_numpy(ref.sph_harm.evaluate_sh_basis(lmax=2, directions=view_dir.reshape(-1, 3).numpy())
                          .reshape(*view_dir.shape[:3], -1).astype(np.float32))
    return dict(uv_map=uv_map, alpha_map=alpha_map, normal_map=normal_map, TBN_map=TBN, view_dir_map=view_dir,
                view_dir_map_tangent=vdt, sh_basis_map=sh)


def _set_bn_train(m):
    if type(m) == torch.nn.BatchNorm2d:
        m.train()


def _compare_png(png_fp, ref_img, alpha_png_fp, ref_alpha, what):
    import cv2
    got = cv2.cvtColor(cv2.imread(png_fp, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
    a_got = cv2.imread(alpha_png_fp, cv2.IMREAD_UNCHANGED).astype(np.float32) / 255.0
    if a_got.ndim == 3:
        a_got = a_got[..., 0]
    agree = (a_got > 0.5) == (ref_alpha.numpy() > 0.5)
    mism = 1.0 - agree.mean()
    want = ref_img.clamp(0, 1).permute(1, 2, 0).numpy()
    p = psnr(torch.from_numpy(got[agree]), torch.from_numpy(want[agree]))
    print('%s: PSNR %.1f dB over %.3f %% of the pixels (coverage differs on %d pixels)' % (what, p, 100 * agree.mean(), int((~agree).sum())))
    assert mism <= 5e-4, '%s: coverage differs on %.4f %% of the pixels' % (what, 100 * mism)
    assert p >= 50.0, '%s: PSNR %.1f dB' % (what, p)
    return p


def test_precompute_outputs_match_the_reference_functions(scene):
    """precompute.py:140-253 through the drop-ins: every map dataio.py:219-245 later loads exists, and view 0's maps equal the
    CPU G-buffer (oracle rasterizer + real reference render / camera functions) -- uv / normal / view-dir max-abs <= 1e-4 on
    agreeing pixels, SH basis <= 1e-5, alpha agreeing on >= 99.95 % of the pixels."""
    import cv2
    import scipy.io
    root = scene['root']
    hi = os.path.join(root, 'precomp_mesh', 'resol_%d' % SIZE)
    for d in ('TBN_map', 'uv_map', 'normal_map', 'view_dir_map', 'view_dir_map_tangent', 'sh_basis_map', 'reflect_dir_map', 'raster'):
        assert len(glob.glob(os.path.join(hi, d, '*.mat'))) == scene['n_views'], d
    assert len(glob.glob(os.path.join(hi, 'alpha_map', '*.png'))) == scene['n_views']
    assert len(glob.glob(os.path.join(root, 'precomp_mesh_7500v', 'resol_%d' % SIZE, 'raster', '*.mat'))) == scene['n_views']
    ref = ref_import.import_reference()
    mesh = _read_obj(os.path.join(root, 'mesh.obj'))
    vw = _views(os.path.join(root, 'calib.mat'), SIZE)[0]
    g = _gbuffer(ref, mesh, vw, SIZE)
    alpha = cv2.imread(os.path.join(hi, 'alpha_map', '00000.png'), cv2.IMREAD_UNCHANGED).astype(np.float32) / 255.0
    agree = (alpha > 0.5) == (g['alpha_map'][0].numpy() > 0.5)
    print('precompute alpha: coverage differs on %d pixels' % int((~agree).sum()))
    assert agree.mean() >= 0.9995
    fg = torch.from_numpy(agree & (alpha > 0.5))
    for name, key, tol in (('uv_map', 'uv_map', 1e-4), ('normal_map', 'normal_map', 1e-4), ('view_dir_map', 'view_dir_map', 1e-5),
                           ('sh_basis_map', 'sh_basis_map', 1e-5), ('view_dir_map_tangent', 'view_dir_map_tangent', 1e-3),
                           ('TBN_map', 'TBN_map', 1e-3)):
        got = torch.from_numpy(scipy.io.loadmat(os.path.join(hi, name, '00000.mat'))[name].astype(np.float32))
        want = g[key][0]
        d = (got - want).abs()
        if name == 'uv_map':
            d = torch.minimum(d, 1 - d)             # u wraps at the seam (uv - floor(uv), network.py:196)
        # the maps are piecewise smooth; compare on the interior of the agreeing foreground (face-index ties at shared edges
        # may pick either neighbour, whose interpolated values agree to the tolerance anyway)
        err = d[fg].max().item()
        q = torch.quantile(d[fg].flatten()[:2000000].float(), 0.999).item()
        print('precompute %s: max-abs %.2e, 99.9 %% quantile %.2e' % (name, err, q))
        assert q <= tol, (name, q)


def test_one_pass_precompute_equals_the_script_output_and_feeds_the_step(scene, tmp_path):
    """relightable_nr_b200.precompute.ViewPrecompute (the whole of precompute.py:140-253 as one device pass, no
    numpy / .mat round trips) produces the maps the UNCHANGED script wrote for the same cameras -- equal to 1e-6 (same kernels; the
    script's copy went through float64 .mat files and an 8-bit alpha PNG) -- and a PackedViewCache of them feeds a training step."""
    import scipy.io
    import cv2
    from relightable_nr_b200.precompute import PackedViewCache, ViewPrecompute
    from relightable_nr_b200.pipeline import RNRPipeline
    root = scene['root']
    hi = os.path.join(root, 'precomp_mesh', 'resol_%d' % SIZE)
    pre = ViewPrecompute(os.path.join(root, 'mesh.obj'), SIZE, global_RT=torch.eye(4))
    vws = _views(os.path.join(root, 'calib.mat'), SIZE)
    proj = torch.cat([v['proj'] for v in vws]).cuda()
    pose = torch.cat([v['pose'] for v in vws]).cuda()
    maps = pre.maps(proj, pose, with_raster=True)
    assert maps['uv_map'].shape == (scene['n_views'], SIZE, SIZE, 2)
    for i in range(scene['n_views']):
        name = '%05d' % i
        for key in ('TBN_map', 'uv_map', 'normal_map', 'view_dir_map', 'view_dir_map_tangent', 'sh_basis_map', 'reflect_dir_map'):
            want = scipy.io.loadmat(os.path.join(hi, key, name + '.mat'))[key].astype(np.float32)
            if key == 'uv_map':
                want = want - np.floor(want)                              # dataio.py:228
            err = np.abs(maps[key][i].cpu().numpy() - want).max()
            assert err <= 1e-6, (key, i, err)
        alpha = cv2.imread(os.path.join(hi, 'alpha_map', name + '.png'), cv2.IMREAD_UNCHANGED).astype(np.float32) / 255.0
        assert np.array_equal(maps['alpha_map'][i].cpu().numpy(), alpha)
    # the reference layout written by the pass itself is readable by the reference's own dataio.ViewDataset
    out_dir = str(tmp_path / 'precomp_mesh')
    pre.write_reference_layout(out_dir, ['%05d' % i for i in range(scene['n_views'])], maps)
    sys.path.insert(0, REF)
    ref_import.import_reference()
    import dataio
    ds = dataio.ViewDataset(root_dir=root, img_dir=os.path.join(root, 'rgb0') + '/', calib_path=os.path.join(root, 'calib.mat'), calib_format='convert',
                            img_size=[SIZE, SIZE], sampling_pattern='all', load_precompute=True, precomp_high_dir=out_dir, precomp_low_dir=out_dir)
    item = ds.read_view(1)
    for key in ('TBN_map', 'uv_map', 'normal_map', 'view_dir_map', 'view_dir_map_tangent', 'sh_basis_map', 'alpha_map'):
        assert np.abs(item[key].numpy().astype(np.float32) - maps[key][1].cpu().numpy()).max() <= 1e-6, key
    # packed cache -> pinned staging -> device -> one fused training step per view
    keys = ('uv_map', 'sh_basis_map', 'normal_map', 'view_dir_map', 'view_dir_map_tangent', 'TBN_map', 'alpha_map')
    views = []
    for i in range(scene['n_views']):
        v = {k: maps[k][i] for k in keys}
        v['img_gt'] = item['img_gt'] if i == 1 else ds.read_view(i)['img_gt']
        views.append(v)
    cache = PackedViewCache.write(str(tmp_path / 'views.rnrcache'), views)
    pipe = RNRPipeline(device='cuda:0', img_size=SIZE, dropout=False)
    copy_stream = torch.cuda.Stream()
    losses = []
    for it in range(4):
        v = cache.load(it % len(cache), device='cuda:0', stream=copy_stream, slot=it)
        torch.cuda.current_stream().wait_stream(copy_stream)
        assert torch.equal(v['uv_map'][0], maps['uv_map'][it % len(cache)])
        losses.append(pipe.train_step(v, fused=True)[0].item())
        torch.cuda.synchronize()
    print('losses from the packed cache:', losses)
    assert all(np.isfinite(losses))


def test_train_rnr_then_test_rnr(scene):
    root = scene['root']
    out = _run('train_rnr.py', '--data_root', root, '--gpu_id', '0', '--img_size', SIZE, '--lp_dir', '_/light_probe',
               '--lighting_relight_idx', '1', '--sphere_samples_fp', '_/sphere_samples_4096.mat', '--max_iter', '4', '--ckp_freq', '2',
               '--log_freq', '2', '--val_freq', '1000', '--exp_name', 'gate')
    iters = [ln for ln in out.splitlines() if ln.startswith('Iter ')]
    assert len(iters) == 4, out[-2000:]
    assert any(ln.startswith('Val   mae_valid') for ln in out.splitlines()), 'the validation pass of iteration 0 must have run'
    print('\n'.join(iters))
    logs = sorted(glob.glob(os.path.join(root, 'logs', 'rnr', '*_gate')))
    assert len(logs) == 1
    ckpts = sorted(glob.glob(os.path.join(logs[0], 'model_epoch-*_iter-*.pth')))
    assert [os.path.basename(c).split('iter-')[1] for c in ckpts] == ['2.pth', '4.pth'], ckpts
    assert os.path.isfile(os.path.join(logs[0], 'params.txt'))
    assert len(glob.glob(os.path.join(logs[0], 'val_out', '*.png'))) == scene['n_views']
    ck = os.path.basename(ckpts[-1])
    _run('test_rnr.py', '--calib_dir', '_/test_seq/spiral_step720', '--checkpoint_dir', logs[0], '--checkpoint_name', ck,
         '--gpu_id', '0', '--img_size', SIZE)
    res = glob.glob(os.path.join(scene['test_calib_dir'], 'resol_%d' % SIZE, 'rnr', '*'))
    assert len(res) == 1, res
    pngs = sorted(glob.glob(os.path.join(res[0], 'img_est_SH_000', '*.png')))
    assert len(pngs) == scene['n_test_views']

    # ---- the same views rendered by the REAL reference modules on the CPU from the checkpoint the drop-ins wrote ----
    ref = ref_import.import_reference()
    import scipy.io
    sd = torch.load(ckpts[-1], map_location='cpu')
    l_dir = torch.from_numpy(scipy.io.loadmat(os.path.join(root, 'sphere_samples_4096.mat'))['sphere_samples'].transpose().copy())
    tm = ref.network.TextureMapper(texture_size=512, texture_num_ch=24, mipmap_level=4, texture_init=None, fix_texture=True, apply_sh=True)
    tm.load_state_dict(sd['texture_mapper'], strict=True)
    rs = ref.network.RaySampler(num_azi=sd['ray_sampler']['num_azi'].numpy(), num_polar=sd['ray_sampler']['num_polar'].numpy(),
                                interval_polar=sd['ray_sampler']['interval_polar'].numpy())
    rs.load_state_dict(sd['ray_sampler'], strict=True)
    rsd = ref.network.RaySampler(num_azi=sd['ray_sampler_diffuse']['num_azi'].numpy(), num_polar=sd['ray_sampler_diffuse']['num_polar'].numpy(),
                                 interval_polar=sd['ray_sampler_diffuse']['interval_polar'].numpy(), mode='diffuse')
    rsd.load_state_dict(sd['ray_sampler_diffuse'], strict=True)
    R = rs.num_ray + rsd.num_ray
    net = ref.network.RenderingNet(nf0=64, in_channels=R * 3 + 6 + 24, out_channels=3 * R, num_down_unet=5, out_channels_gcn=512)
    net.load_state_dict(sd['render_net'], strict=True)                       # the drop-in's checkpoint in the real class
    import cv2
    probes = []
    for fp in sorted(glob.glob(os.path.join(root, 'light_probe', '*.png'))):
        im = cv2.imread(fp, cv2.IMREAD_UNCHANGED)[:, :, :3].astype(np.float32) / 255.0
        probes.append({'lp_img': torch.from_numpy(cv2.cvtColor(im, cv2.COLOR_BGR2RGB).transpose(2, 0, 1))[None]})
    lm_lp = ref.network.LightingLP(l_dir, num_channel=3, lp_dataloader=probes, fix_params=True)
    lm_lp.fit_sh(lmax=10)
    lm = ref.network.LightingSH(l_dir, lmax=10, num_lighting=lm_lp.num_lighting, num_channel=3, init_coeff=lm_lp.sh_coeff, fix_params=True)
    rr = ref.network.RayRenderer(lm, ref.network.Interpolater())
    for m in (tm, rs, rsd, net, rr, lm):
        m.eval()
    net.apply(_set_bn_train)
    mesh = _read_obj(os.path.join(root, 'mesh.obj'))
    worst = 1e9
    with torch.no_grad():
        for i, vw in enumerate(_views(os.path.join(scene['test_calib_dir'], 'calib.mat'), SIZE)):
            g = _gbuffer(ref, mesh, vw, SIZE)
            a = g['alpha_map'][..., None]
            neural = tm(g['uv_map'], g['sh_basis_map'], sh_start_ch=6)
            d0, uv0, _ = rs(g['TBN_map'], g['view_dir_map_tangent'], a)
            d1, uv1, _ = rsd(g['TBN_map'], g['view_dir_map_tangent'], a)
            rays_dir, rays_uv = torch.cat((d0, d1), -1), torch.cat((uv0, uv1), -1)
            x = torch.cat((rays_dir.permute(0, -1, -2, 1, 2).reshape(1, -1, SIZE, SIZE), g['normal_map'].permute(0, 3, 1, 2),
                           g['view_dir_map'].permute(0, 3, 1, 2), neural), 1)
            lt = (net(x, sd['v_feature']).reshape(1, R, -1, SIZE, SIZE) * 0.5 + 0.5) * 2.0
            out = rr(neural[:, 3:6], rays_uv, lt, lighting_idx=0, albedo_diffuse=neural[:, :3], num_ray_diffuse=d1.shape[-1],
                     lp_scale_factor=1, seperate_albedo=True)[0]
            worst = min(worst, _compare_png(pngs[i], out[0], os.path.join(res[0], 'alpha_map', '%05d.png' % i), g['alpha_map'][0],
                                            'test_rnr.py view %d vs the real reference modules' % i))
    print('test_rnr.py renders vs real reference: worst PSNR %.1f dB' % worst)


def test_train_dnr_then_test_dnr(scene):
    root = scene['root']
    out = _run('train_dnr.py', '--data_root', root, '--gpu_id', '0', '--img_size', SIZE, '--texture_num_ch', '16', '--max_epoch', '1',
               '--ckp_freq', '2', '--log_freq', '2', '--exp_name', 'gate')
    iters = [ln for ln in out.splitlines() if ln.startswith('Iter ')]
    assert len(iters) == scene['n_views'], out[-2000:]
    print('\n'.join(iters))
    logs = sorted(glob.glob(os.path.join(root, 'logs', 'dnr', '*_gate')))
    assert len(logs) == 1
    ckpts = sorted(glob.glob(os.path.join(logs[0], 'model_epoch-*_iter-*.pth')))
    assert ckpts, os.listdir(logs[0])
    ck = os.path.basename(ckpts[-1])
    _run('test_dnr.py', '--calib_dir', '_/test_seq/spiral_step720', '--checkpoint_dir', logs[0], '--checkpoint_name', ck,
         '--gpu_id', '0', '--img_size', SIZE, '--force_recompute', '1')
    res = [d for d in glob.glob(os.path.join(scene['test_calib_dir'], 'resol_%d' % SIZE, 'dnr', '*')) if os.path.isdir(os.path.join(d, 'img_est'))]
    assert len(res) == 1, res
    pngs = sorted(glob.glob(os.path.join(res[0], 'img_est', '*.png')))
    assert len(pngs) == scene['n_test_views']
    ref = ref_import.import_reference()
    sd = torch.load(ckpts[-1], map_location='cpu')
    tm = ref.network.TextureMapper(texture_size=512, texture_num_ch=16, mipmap_level=4, texture_init=None, fix_texture=True, apply_sh=True)
    tm.load_state_dict(sd['texture_mapper'], strict=True)
    net = ref.network.RenderingNet(nf0=80, in_channels=16, out_channels=3, num_down_unet=5, use_gcn=False)
    net.load_state_dict(sd['render_net'], strict=True)
    tm.eval()
    net.eval()
    net.apply(_set_bn_train)
    mesh = _read_obj(os.path.join(root, 'mesh.obj'))
    with torch.no_grad():
        for i, vw in enumerate(_views(os.path.join(scene['test_calib_dir'], 'calib.mat'), SIZE)):
            g = _gbuffer(ref, mesh, vw, SIZE)
            out = (net(tm(g['uv_map'], g['sh_basis_map']), None) * 0.5 + 0.5) * 2.0 * g['alpha_map'][:, None]
            _compare_png(pngs[i], out[0], os.path.join(res[0], 'alpha_map', '%05d.png' % i), g['alpha_map'][0],
                         'test_dnr.py view %d vs the real reference modules' % i)
