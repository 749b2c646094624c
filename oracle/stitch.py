"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the per-view scatter of stitch_lp.py -- rays, spherical
mapping, fancy-index accumulation, final division -- used to check csrc/stitch.cu on the GPU box, where /root/reference does not
exist.  Pinned against the UNCHANGED stitch_lp.py run through the launcher in tests/test_stitch.py (CPU test).

  view_rays           stitch_lp.py:27-35   camera2ray
  probe_texels        stitch_lp.py:22-24   spherical_mapping, :139-143 scale / clip / round
  scatter_view        stitch_lp.py:145-146 numpy fancy-index "+=" (the last duplicate wins; one count per texel and view)
  finish              stitch_lp.py:149-150
"""
import numpy as np


def view_rays(pose, proj, w, h):
    ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing='ij')
    p = np.stack((xs, ys, np.ones_like(xs)), 0).reshape(3, -1)
    d = np.linalg.inv(pose[:3, :3]).dot(np.linalg.inv(proj).dot(p))
    d = d / np.sqrt((d * d).sum(0))[None]
    return d.reshape(3, h, w)


def probe_texels(dirs, lp_h, lp_w):
    """dirs [3, n] unit vectors -> (row, column) integer probe coordinates."""
    u = np.arctan2(dirs[2], dirs[0]) * 0.5 / np.pi + 0.5
    v = np.arccos(dirs[1]) * 1.0 / np.pi
    u = (u * lp_w).clip(max=lp_w - 1.0)
    v = (v * lp_h).clip(max=lp_h - 1.0)
    return np.round(v).astype('int'), np.round(u).astype('int')


def scatter_view(env, count, img, bg_mask, pose, proj):
    h, w = bg_mask.shape
    rows, cols = probe_texels(view_rays(pose, proj, w, h)[:, bg_mask], env.shape[0], env.shape[1])
    env[rows, cols] += img[bg_mask][:, :3]
    count[rows, cols] += 1


def finish(env, count):
    mask = count.sum(2) > 0
    env[mask] /= count[mask]
    return env, mask
