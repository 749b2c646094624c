#!/usr/bin/env python
"""Synthetic scene in the on-disk layout the reference's UNCHANGED scripts read (README.md:40-53,
dataio.py:48-50,219-245, train_rnr.py:183-339, test_rnr.py:50-103): the material-sphere proxy.

    python tools/make_scene.py /tmp/scene --views 4 --test-views 3 --img-size 512

writes
    calib.mat                              poses [n,4,4] world->camera, projs [n,3,3], img_hws [n,2], dist_coeffs [n,5], global_RT [4,4]
    mesh.obj                               unit UV sphere, lat x lon quads split in two, seam-duplicated v / vt / vn  (f v/vt/vn)
    mesh_7500v.obj                         75 x 100 grid sphere = exactly 7500 vertices (GCN input, train_rnr.py:257-258)
    tex.png                                all-white 512^2 (README.md:52)
    rgb0/<name>.png                        n training views of the analytically shaded sphere
    light_probe/{0,1}.png                  two equirect probes (sorted order = lighting index)
    light_probe_stitch_all/0.png, mask/0.png, count/0.mat     stitched probe of the training lighting (train_rnr.py:293-305)
    sphere_samples_4096.mat                the reference's own quadrature directions (from tests/golden/sphere_samples_4096.npz)
    test_seq/spiral_step720/calib.mat      inference cameras (test_rnr.py:97-103)
The per-view precomputed maps (precomp_mesh/..., precomp_mesh_7500v/...) are NOT written here: they are produced by running
the unchanged precompute.py on this scene through the launcher.  No reference code is involved; pure numpy / OpenCV / scipy.
"""
import argparse
import math
import os

import numpy as np


def spiral_pose(i, radius=3.0):
    """Camera i of spiral_step720 (camera.py:72-76: azimuth -2 deg, elevation 0.125 deg per step) looking at the origin:
    4x4 world->camera matrix with x right, y down, z forward (camera.py:48-69) and the camera position."""
    azi, ele = math.radians(-2.0 * i), math.radians(0.125 * i)
    pos = np.array([radius * math.cos(ele) * math.sin(azi), radius * math.sin(ele), radius * math.cos(ele) * math.cos(azi)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    RT = np.eye(4)
    RT[:3, :3] = np.stack((right, -up, fwd))
    RT[:3, 3] = -RT[:3, :3].dot(pos)
    return RT, pos


def uv_sphere(n_lat, n_lon):
    """v [nv,3], vt [nv,2], vn [nv,3], f [nf,3] (0-based, one index for v/vt/vn); vt = (lon/2pi, 1 - lat/pi), vn = v."""
    lat = np.linspace(0.0, np.pi, n_lat + 1)
    lon = np.linspace(0.0, 2 * np.pi, n_lon + 1)
    la, lo = np.meshgrid(lat, lon, indexing='ij')
    v = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    vt = np.stack([lo / (2 * np.pi), 1 - la / np.pi], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n_lat), np.arange(n_lon), indexing='ij')
    a = (i * (n_lon + 1) + j).reshape(-1)
    b, c, d = a + 1, a + n_lon + 1, a + n_lon + 2
    f = np.stack([np.stack([a, b, c], 1), np.stack([b, d, c], 1)], 1).reshape(-1, 3)
    return v, vt, v.copy(), f


def grid_sphere(n_lat, n_lon):
    """n_lat x n_lon vertices (no seam / pole duplicates: rings strictly between the poles, wrap-around in longitude)."""
    lat = (np.arange(n_lat) + 0.5) / n_lat * np.pi
    lon = np.arange(n_lon) / n_lon * 2 * np.pi
    la, lo = np.meshgrid(lat, lon, indexing='ij')
    v = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    vt = np.stack([lo / (2 * np.pi), 1 - la / np.pi], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n_lat - 1), np.arange(n_lon), indexing='ij')
    a = (i * n_lon + j).reshape(-1)
    b = (i * n_lon + (j + 1) % n_lon).reshape(-1)
    c, d = a + n_lon, b + n_lon
    f = np.stack([np.stack([a, b, c], 1), np.stack([b, d, c], 1)], 1).reshape(-1, 3)
    return v, vt, v.copy(), f


def write_obj(path, v, vt, vn, f):
    with open(path, 'w') as fh:
        fh.write('# synthetic material-sphere proxy (tools/make_scene.py)\n')
        fh.write(''.join('v %.8f %.8f %.8f\n' % tuple(p) for p in v))
        fh.write(''.join('vt %.8f %.8f\n' % tuple(p) for p in vt))
        fh.write(''.join('vn %.8f %.8f %.8f\n' % tuple(p) for p in vn))
        fh.write(''.join('f %d/%d/%d %d/%d/%d %d/%d/%d\n' % (t[0], t[0], t[0], t[1], t[1], t[1], t[2], t[2], t[2]) for t in f + 1))


def envmap(h, w, which):
    """Smooth equirect probe [h,w,3] in [0,1]: sky gradient + one warm / cool lobe (distinct per lighting index)."""
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing='ij')
    sky = 0.25 + 0.35 * (1 - v)
    cu, cv = (0.3, 0.35) if which == 0 else (0.7, 0.45)
    lobe = np.exp(-(((u - cu + 0.5) % 1.0 - 0.5) ** 2 / 0.02 + (v - cv) ** 2 / 0.03))
    tint = np.array([1.0, 0.85, 0.6]) if which == 0 else np.array([0.6, 0.8, 1.0])
    return np.clip(sky[..., None] * np.array([0.8, 0.9, 1.0]) + 0.6 * lobe[..., None] * tint, 0, 1)


def shade_sphere(RT, pos, K, size, light):
    """Analytic image of the unit sphere from camera (RT, K): Lambert + a little specular from direction ``light``."""
    v, u = np.meshgrid(np.arange(size) + 0.5, np.arange(size) + 0.5, indexing='ij')
    cam = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)
    d = cam @ RT[:3, :3]                      # R^T applied to row vectors
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    b = d @ pos
    disc = b * b - (pos @ pos - 1.0)
    hit = disc > 0
    t = -b - np.sqrt(np.maximum(disc, 0))
    n = pos + t[..., None] * d
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-9)
    lam = np.maximum(n @ light, 0.0)
    hvec = light - d
    hvec /= np.maximum(np.linalg.norm(hvec, axis=-1, keepdims=True), 1e-9)
    spec = np.maximum((n * hvec).sum(-1), 0.0) ** 40
    albedo = 0.55 + 0.25 * np.stack([np.sin(5 * n[..., 0]), np.sin(5 * n[..., 1] + 1), np.sin(5 * n[..., 2] + 2)], -1)
    img = albedo * (0.25 + 0.75 * lam[..., None]) + 0.3 * spec[..., None]
    return np.clip(img, 0, 1) * hit[..., None]


def write_calib(path, idxs, size, radius=3.0):
    import scipy.io
    n = len(idxs)
    K = np.array([[1.2 * size, 0, size / 2.0], [0, 1.2 * size, size / 2.0], [0, 0, 1.0]])
    poses = np.stack([spiral_pose(i, radius)[0] for i in idxs])
    scipy.io.savemat(path, {'poses': poses, 'projs': np.repeat(K[None], n, 0), 'img_hws': np.full((n, 2), size, dtype=np.int64),
                            'dist_coeffs': np.zeros((n, 5)), 'global_RT': np.eye(4)})
    return K, poses


def make_scene(root, n_views=4, n_test_views=3, img_size=512, mesh_lat=128, mesh_lon=256, probe_hw=(64, 128), view_stride=7, seed=0):
    """Write the scene under ``root``; returns a dict of the paths a caller needs."""
    import cv2
    import scipy.io
    os.makedirs(root, exist_ok=True)
    for d in ('rgb0', 'light_probe', 'light_probe_stitch_all/mask', 'light_probe_stitch_all/count', 'test_seq/spiral_step720'):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    write_obj(os.path.join(root, 'mesh.obj'), *uv_sphere(mesh_lat, mesh_lon))
    write_obj(os.path.join(root, 'mesh_7500v.obj'), *grid_sphere(75, 100))
    cv2.imwrite(os.path.join(root, 'tex.png'), np.full((512, 512, 3), 255, np.uint8))
    train_idx = [view_stride * i for i in range(n_views)]
    K, poses = write_calib(os.path.join(root, 'calib.mat'), train_idx, img_size)
    write_calib(os.path.join(root, 'test_seq', 'spiral_step720', 'calib.mat'), [3 + 11 * i for i in range(n_test_views)], img_size)
    light = np.array([0.5, 0.6, 0.62])
    light /= np.linalg.norm(light)
    for k, i in enumerate(train_idx):
        RT, pos = spiral_pose(i)
        img = shade_sphere(RT, pos, K, img_size, light)
        cv2.imwrite(os.path.join(root, 'rgb0', '%05d.png' % k), (img[:, :, ::-1] * 255 + 0.5).astype(np.uint8))
    ph, pw = probe_hw
    for which in (0, 1):
        cv2.imwrite(os.path.join(root, 'light_probe', '%d.png' % which), (envmap(ph, pw, which)[:, :, ::-1] * 255 + 0.5).astype(np.uint8))
    # stitched probe of lighting 0: the probe itself where "observed", with an unobserved band (mask 0) near the bottom
    rng = np.random.RandomState(seed)
    st = np.clip(envmap(ph, pw, 0) + 0.02 * rng.randn(ph, pw, 3), 0, 1)
    mask = np.ones((ph, pw, 3), np.uint8) * 255
    mask[int(0.8 * ph):, :, :] = 0
    st[mask == 0] = 0
    cv2.imwrite(os.path.join(root, 'light_probe_stitch_all', '0.png'), (st[:, :, ::-1] * 255 + 0.5).astype(np.uint8))
    cv2.imwrite(os.path.join(root, 'light_probe_stitch_all', 'mask', '0.png'), mask)
    count = (mask[:, :, 0] > 0).astype(np.float64) * rng.randint(1, n_views + 1, size=(ph, pw))
    scipy.io.savemat(os.path.join(root, 'light_probe_stitch_all', 'count', '0.mat'), {'count': count, 'num_view': np.array([[n_views]], dtype=np.float64)})
    samples = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'sphere_samples_4096.npz'))
    scipy.io.savemat(os.path.join(root, 'sphere_samples_4096.mat'), {'sphere_samples': samples['sphere_samples']})
    return {'root': root, 'n_views': n_views, 'n_test_views': n_test_views, 'img_size': img_size,
            'test_calib_dir': os.path.join(root, 'test_seq', 'spiral_step720')}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('root')
    ap.add_argument('--views', type=int, default=4)
    ap.add_argument('--test-views', type=int, default=3)
    ap.add_argument('--img-size', type=int, default=512)
    ap.add_argument('--mesh-lat', type=int, default=128)
    ap.add_argument('--mesh-lon', type=int, default=256)
    a = ap.parse_args()
    print(make_scene(a.root, a.views, a.test_views, a.img_size, a.mesh_lat, a.mesh_lon))
