"""GPU parity of the general-degree SH basis kernel (rnr_sh_basis -> sph_harm.evaluate_sh_basis, a9 / a17): the lmax-10 tables
behind LightingSH.basis_val / basis_val_recon (network.py:557,581) and LightingLP.fit_sh (:696).

The oracle (oracle.pixel_ops.evaluate_sh_basis) is itself pinned, function by function (sign, order, normalisation), to an
independent scipy evaluation in tests/test_oracle_golden.py.  Two gates:
  * exact-math gate: kernel vs the oracle fed the SAME float32 directions promoted to float64 -- <= 1e-9 (both fp64);
  * reference-rounding gate: kernel vs the oracle fed float32 directions, which then rounds azimuth / colatitude to float32
    DEGREES exactly like sph_harm.py:54-57 does -- <= 2e-5: that rounding (<= 2 ulp of 180 deg = 5e-7 rad) times
    max |dY/d angle| (~ 20 at l = 10) is noise of the reference itself, not of the kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import pixel_ops as P

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def _grid_dirs(h, w):
    vv, uu = torch.meshgrid(torch.arange(h, dtype=torch.float32) / (h - 1), torch.arange(w, dtype=torch.float32) / (w - 1), indexing='ij')
    return P.spherical_mapping_inv(torch.stack((uu, vv)).flatten(1)).t().contiguous().numpy()


@pytest.mark.parametrize('which', ['sphere_samples_4096', 'grid_256x512'])
@pytest.mark.parametrize('lmax', [10, 4])
def test_sh_basis_general_kernel_vs_pinned_oracle(which, lmax):
    from relightable_nr_b200.dropin import sph_harm
    if which == 'sphere_samples_4096':
        d = np.load(os.path.join(G, 'sphere_samples_4096.npz'))['sphere_samples'].astype(np.float32)
    else:
        d = _grid_dirs(256, 512).astype(np.float32)
    Y = sph_harm.evaluate_sh_basis(lmax=lmax, directions=d)
    assert Y.shape == (d.shape[0], (lmax + 1) ** 2) and Y.dtype == np.float64
    exact = P.evaluate_sh_basis(lmax, d.astype(np.float64))
    e1 = np.abs(Y - exact).max()
    ref_rounding = P.evaluate_sh_basis(lmax, d)
    e2 = np.abs(Y - ref_rounding).max()
    print('%s lmax %d: max-abs vs exact-math oracle %.2e, vs float32-degree (reference rounding) oracle %.2e' % (which, lmax, e1, e2))
    assert e1 <= 1e-9
    assert e2 <= 2e-5


def test_sh_basis_azi_pol_entry_and_lighting_tables():
    """The azi / pol (degrees) entry of sph_harm.evaluate_sh_basis, and the two LightingSH tables built through it."""
    from relightable_nr_b200.dropin import network, sph_harm
    rng = np.random.RandomState(3)
    azi, pol = rng.uniform(-180, 180, 300), rng.uniform(0, 180, 300)
    Y = sph_harm.evaluate_sh_basis(lmax=6, azi=azi, pol=pol)
    a, p = np.deg2rad(azi), np.deg2rad(pol)
    d = np.stack((np.sin(p) * np.cos(a), np.sin(p) * np.sin(a), np.cos(p)), 1)
    assert np.abs(Y - P.evaluate_sh_basis(6, d)).max() <= 2e-6      # directions pass through float32 on their way to the kernel
    l_dir = torch.from_numpy(np.load(os.path.join(G, 'sphere_samples_4096.npz'))['sphere_samples']).t().contiguous()
    lm = network.LightingSH(l_dir.cuda(), lmax=10, num_lighting=1, lp_recon_h=32, lp_recon_w=64)
    ref = torch.from_numpy(P.evaluate_sh_basis(10, l_dir.t().numpy())).float()
    assert (lm.basis_val.cpu() - ref).abs().max().item() <= 2e-5
    ref = torch.from_numpy(P.evaluate_sh_basis(10, _grid_dirs(32, 64))).float()
    assert (lm.basis_val_recon.cpu() - ref).abs().max().item() <= 2e-5
