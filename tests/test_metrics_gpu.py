"""relightable_nr_b200.metrics.compute_err_metrics_batch (device) against the reference's own metric.py (numpy, host) from the staged
copy: every key of metric.py:87-113 to 1e-5 relative (SSIM through the same Gaussian-window implementation on both sides:
pytorch_msssim itself is not installable here -- relightable_nr_b200/compat/pytorch_msssim.py)."""
import sys

import numpy as np
import pytest
import torch

from tests.golden import ref_import

pytestmark = pytest.mark.gpu


def test_device_metrics_match_reference_metric_py():
    if not ref_import.available():
        pytest.skip('reference not staged')
    from relightable_nr_b200.compat import pytorch_msssim as shim
    sys.modules.setdefault('pytorch_msssim', shim)
    if ref_import.REF not in sys.path:
        sys.path.insert(0, ref_import.REF)
    import metric as ref_metric
    from relightable_nr_b200 import metrics
    g = torch.Generator().manual_seed(0)
    N, H, W = 2, 96, 80
    est = torch.rand(N, 3, H, W, generator=g) * 255
    gt = (est + 12 * torch.randn(N, 3, H, W, generator=g)).clamp(0, 255)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    mask = torch.stack([((yy - 40) ** 2 + (xx - 30) ** 2 < 25 ** 2).float(), ((yy > 10) & (yy < 70) & (xx > 20) & (xx < 75)).float()])[:, None]
    got = metrics.compute_err_metrics_batch(est.cuda(), gt.cuda(), mask.cuda(), compute_ssim=True)
    want = ref_metric.compute_err_metrics_batch(est.clone(), gt.clone(), mask.clone(), compute_ssim=True)
    for k, v in want.items():
        assert k in got, k
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(v, dtype=np.float64)
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6), (k, a, b)
    got2 = metrics.compute_err_metrics_batch(est.cuda(), gt.cuda(), mask.cuda(), compute_ssim=False)
    want2 = ref_metric.compute_err_metrics_batch(est.clone(), gt.clone(), mask.clone(), compute_ssim=False)
    assert np.isnan(got2['ssim_mean']) and np.isnan(want2['ssim_mean'])
    assert np.allclose(got2['psnr_valid'], want2['psnr_valid'], rtol=1e-6)
