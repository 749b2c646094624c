"""Per-view geometry operators against fixtures produced by the REAL reference (tests/golden/make_golden.py::gen_geometry):
render.get_TBN_map (a7), camera.get_view_dir_map (a8), network.LightingLP construction (a18).  Tolerance: max-abs <= 1e-5
(1e-4 relative for the HDR probe samples)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geometry.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_tbn_map_matches_reference(g):
    from relightable_nr_b200.dropin import render
    tbn = render.get_TBN_map(g['tbn_normal'].cuda(), g['tbn_fidx'].cuda(), g['tbn_faces_v'].cuda(), g['tbn_faces_vt'].cuda())
    assert tbn.shape == g['tbn_out'].shape
    assert (tbn.cpu() - g['tbn_out']).abs().max().item() <= 1e-5


def test_view_dir_map_matches_reference(g):
    from relightable_nr_b200.dropin import camera
    vd, vdc = camera.get_view_dir_map((9, 11), g['vd_Kinv'].cuda(), g['vd_Rinv'].cuda())
    assert (vd.cpu() - g['vd_out']).abs().max().item() <= 1e-5
    assert (vdc.cpu() - g['vd_cam']).abs().max().item() <= 1e-5


def test_lighting_lp_matches_reference(g):
    from relightable_nr_b200.dropin import network
    probes = [{'lp_img': g['lp_probe0']}, {'lp_img': g['lp_probe1']}]
    lp = network.LightingLP(g['lp_l_dir'].cuda(), lp_dataloader=probes, lp_img_h=20, lp_img_w=40)
    assert lp.num_lighting == 2
    assert (lp.lps.cpu() - g['lp_lps']).abs().max().item() <= 1e-5
    assert (lp.l_samples_uv.cpu() - g['lp_uv']).abs().max().item() <= 1e-5
    ref = g['lp_l_samples']
    assert (lp.l_samples.data.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    out = lp(lighting_idx=1)
    assert out.shape == (1, 64, 3)
    lp.fit_sh(lmax=2)
    assert lp.sh_coeff.shape == (2, 9, 3) and torch.isfinite(lp.sh_coeff).all()
