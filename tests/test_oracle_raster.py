"""Pins oracle/raster.py (CPU restatement of the reference's rasterization path) to
  (a) tests/golden/raster.npz -- projection / vertices_to_faces outputs of the REAL reference Python and z-buffers produced by the
      reference's own CUDA kernel bodies compiled for the CPU (tests/golden/make_golden.py), and
  (b) the live oracle/_ref/libref_raster.so when it is present,
plus known-answer tests of the rasterization rules (single triangle, shared edge, depth tie, back face, near/far).
CPU only.  Integer maps must be bit-exact; float maps are compared exactly too (same operation order, no FMA)."""
import os

import numpy as np
import pytest
import torch

from oracle import raster as Rr
from oracle import ref_raster

G = os.path.join(os.path.dirname(__file__), 'golden', 'raster.npz')


@pytest.fixture(scope='module')
def g():
    z = np.load(G)
    return {k: z[k] for k in z.files}


def test_projection_matches_reference(g):
    t = lambda k: torch.from_numpy(g[k])
    out = Rr.projection(t('pj_v'), t('pj_K'), t('pj_R'), t('pj_t'), t('pj_dist'), 64, t('pj_off'), t('pj_sc'))
    assert torch.allclose(out, t('pj_out'), rtol=1e-6, atol=1e-6)
    out = Rr.projection(t('pj_v'), t('pj_K'), t('pj_R'), t('pj_t'), torch.zeros(2, 5), 64)
    assert torch.allclose(out, t('pj_out_plain'), rtol=1e-6, atol=1e-6)


def test_vertices_to_faces_matches_reference(g):
    t = lambda k: torch.from_numpy(g[k])
    assert torch.equal(Rr.vertices_to_faces(t('pj_out'), t('vf_faces')), t('vf_out'))
    assert torch.equal(Rr.vertex_attrs_to_faces(t('vf_attr_in'), t('vf_faces')), t('vf_attr'))


@pytest.mark.parametrize('scene', ['sphere', 'soup'])
def test_zbuffer_matches_reference_kernels(g, scene):
    size = int(g['zb_size'])
    faces = g['zb_%s_faces' % scene]
    fim, wm, dm, fiv = Rr.face_index_map(faces, size, 0.0, 1e5, return_face_inv=True)
    assert np.array_equal(fim, g['zb_%s_fim' % scene])
    assert np.array_equal(wm, g['zb_%s_wm' % scene])
    assert np.array_equal(dm, g['zb_%s_dm' % scene])
    inv, _ = Rr.face_inv(faces.reshape(1, -1, 9), size)
    assert np.array_equal(np.nan_to_num(inv, nan=0.0, posinf=0.0, neginf=0.0), g['zb_%s_finv' % scene])
    # the bounding-box shortcut of the oracle changes nothing
    b = Rr.face_index_map(faces, size, 0.0, 1e5, brute_force=True)
    assert np.array_equal(b[0], fim) and np.array_equal(b[1], wm) and np.array_equal(b[2], dm)


@pytest.mark.skipif(not ref_raster.available(), reason='oracle/_ref/libref_raster.so not built')
def test_zbuffer_matches_live_reference_kernels():
    rng = np.random.RandomState(3)
    for size, nf in ((33, 40), (64, 300)):
        faces = (rng.rand(2, nf, 3, 3).astype(np.float32) * 2 - 1)
        faces[..., :2] *= 1.3                                   # some triangles leave the image
        faces[..., 2] = faces[..., 2] * 0.4 + 1.0
        faces[:, 5] = faces[:, 4]                               # depth ties
        fim, wm, dm, fiv = Rr.face_index_map(faces, size, 0.0, 1e5, return_face_inv=True)
        r = ref_raster.forward_face_index_map(faces, size, 0.0, 1e5)
        assert np.array_equal(fim, r[0]) and np.array_equal(wm, r[1]) and np.array_equal(dm, r[2]) and np.array_equal(fiv, r[3])


# ---- known answers ------------------------------------------------------------------------------------------------------
def _tri(pts, z=1.0):
    return np.array([[[x, y, z] for x, y in pts]], dtype=np.float32)[None]          # [1,1,3,3]


def test_single_triangle_and_backface():
    size = 8
    ccw = [(-0.9, -0.9), (0.9, -0.9), (-0.9, 0.9)]
    fim, wm, dm = Rr.face_index_map(_tri(ccw, 2.0), size)
    assert (fim >= 0).sum() > 0
    assert np.allclose(dm[fim >= 0], 2.0) and np.allclose(dm[fim < 0], 1e5)
    assert np.allclose(wm[fim >= 0].sum(-1), 1.0, atol=1e-6) and (wm[fim < 0] == 0).all()
    # pixel centres strictly inside the lower-left half only
    yy, xx = np.nonzero(fim[0] >= 0)
    assert ((xx + yy) <= size - 1).all()
    # the same triangle with opposite winding is a back face: nothing drawn (rasterize_cuda_kernel.cu:109)
    fim2, _, _ = Rr.face_index_map(_tri(ccw[::-1], 2.0), size)
    assert (fim2 == -1).all()


def test_shared_edge_has_no_gap_and_depth_tie_prefers_lowest_index():
    size = 16
    a = [(-1.0, -1.0), (1.0, -1.0), (-1.0, 1.0)]
    b = [(1.0, -1.0), (1.0, 1.0), (-1.0, 1.0)]
    faces = np.concatenate([_tri(a), _tri(b)], axis=1)
    fim, _, _ = Rr.face_index_map(faces, size)
    assert (fim >= 0).all()                                     # the two triangles tile the image: no uncovered pixel
    faces = np.concatenate([_tri(a), _tri(a)], axis=1)           # coincident faces: strict '<' keeps face 0 (:142)
    fim, _, _ = Rr.face_index_map(faces, size)
    assert set(np.unique(fim)) <= {-1, 0}


def test_near_far_and_depth_order():
    size = 8
    full = [(-3.0, -3.0), (3.0, -3.0), (0.0, 3.0)]
    faces = np.concatenate([_tri(full, 5.0), _tri(full, 2.0), _tri(full, -1.0), _tri(full, 2e5)], axis=1)
    fim, _, dm = Rr.face_index_map(faces, size, near=0.0, far=1e5)
    assert (fim == 1).all() and np.allclose(dm, 2.0)            # nearest valid face wins; z <= near and z >= far are skipped
    fim, _, _ = Rr.face_index_map(faces, size, near=3.0, far=1e5)
    assert (fim == 0).all()


def test_rasterizer_forward_background_is_exact_zero_and_flipped(g):
    m = dict(vertices=torch.from_numpy(g['zb_mesh_v'])[None], faces=torch.from_numpy(g['zb_mesh_f'])[None],
             vertices_texcoords=torch.from_numpy(g['zb_mesh_vt'])[None], faces_vt_idx=torch.from_numpy(g['zb_mesh_f'])[None],
             vertices_normals=torch.from_numpy(g['zb_mesh_vn'])[None], faces_vn_idx=torch.from_numpy(g['zb_mesh_f'])[None])
    size = int(g['zb_size'])
    out = Rr.rasterizer_forward(m, size, torch.from_numpy(g['zb_K']), torch.from_numpy(g['zb_pose']))
    uv, alpha, fim, w = out[0], out[1], out[2], out[3]
    assert np.array_equal(fim.numpy()[:, ::-1], g['zb_sphere_fim'])          # rasterize.py:313-321 vertical flip
    bg = alpha[0] == 0
    assert bg.any() and (uv[0][bg] == 0).all() and (out[5][0][bg] == 0).all() and (out[9][0][bg] == 0).all()
    fg = ~bg
    assert torch.allclose(out[5][0][fg].norm(dim=-1), torch.ones(int(fg.sum())), atol=1e-5)
    # perspective-correct weights still sum to ~1 on the surface, positions lie on the unit sphere (flat faces: slightly inside)
    assert torch.allclose(w[0][fg].sum(dim=(-1, -2)), torch.ones(int(fg.sum())), atol=1e-3)
    r = out[9][0][fg].norm(dim=-1)
    assert (r <= 1.0 + 1e-5).all() and (r > 0.9).all()
