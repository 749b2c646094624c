"""Validation metrics on the device: drop-in for the reference's ``metric.compute_err_metrics_batch``
(metric.py:19-113), which pulls the rendered image, the ground truth and the mask to the host for every view -- per training
iteration (train_rnr.py:626-633) and per validation view (:707-887) -- and reduces them in numpy.

Here the masked MAE / MSE / PSNR family comes from two kernels (csrc/metrics.cu) and ONE read-back of 8 doubles + a bounding box
per image; SSIM (metric.py:78-84) runs on the device tensors through the Gaussian-window implementation of
relightable_nr_b200/compat/pytorch_msssim.py.  Same keys, same shapes ([N,1] arrays + ``*_mean`` scalars), same conventions:
images on a 0..255 scale, pixels with mask != 1 zeroed in both images, PSNR = 100 when the mse is < 1e-10 (metric.py:7-16)."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .compat.pytorch_msssim import ssim as _ssim

vp, i32 = C.c_void_p, C.c_int
_lib.register_sigs({"rnr_metric_sums": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]})

_KEYS = ('mae', 'mae_bb', 'mae_valid', 'mse', 'mse_bb', 'mse_valid', 'psnr', 'psnr_bb', 'psnr_valid', 'ssim', 'ssim_bb', 'ssim_valid')


def _psnr(mse255):
    mse = mse255 / (255.0 * 255.0)
    return 100 if mse < 1.0e-10 else 20 * math.log10(1.0 / math.sqrt(mse))


def compute_err_metrics_batch(img_est, img_gt, mask, compute_ssim=True):
    """img_est, img_gt [N,3,H,W] (0..255), mask [N,1,H,W] -> dict as metric.compute_err_metrics_batch (metric.py:87-113)."""
    if not (img_est.is_cuda and img_gt.is_cuda and mask.is_cuda):
        raise TypeError('metrics.compute_err_metrics_batch: CUDA tensors required (librnr_b200 has no CPU path)')
    N, Cc, H, W = img_est.shape
    m = (mask.reshape(N, 1, H, W) == 1).float().contiguous()
    est = (img_est.float() * m).contiguous()
    gt = (img_gt.float() * m).contiguous()
    dev = est.device
    box = torch.tensor([W, -1, H, -1], dtype=torch.int32, device=dev).repeat(N, 1).contiguous()
    cnt = torch.zeros(N, dtype=torch.int64, device=dev)
    sums = torch.zeros((N, 4), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().rnr_metric_sums(est.data_ptr(), gt.data_ptr(), m.data_ptr(), N, Cc, H, W, box.data_ptr(), cnt.data_ptr(),
                                          sums.data_ptr(), torch.cuda.current_stream().cuda_stream), 'rnr_metric_sums')
    packed = torch.cat((sums, box.double(), cnt.double()[:, None]), 1).cpu().numpy()       # the one read-back
    out = {k: [] for k in _KEYS}
    for n in range(N):
        s_abs, s_sq, s_abs_bb, s_sq_bb, xmin, xmax, ymin, ymax, c = packed[n]
        if c == 0:
            raise ValueError('compute_err_metrics: empty mask (the reference fails on min() of an empty sequence, metric.py:48)')
        xmin, xmax, ymin, ymax = int(xmin), int(xmax) + 1, int(ymin), int(ymax) + 1
        n_all, n_bb, n_valid = float(H * W * Cc), float((ymax - ymin) * (xmax - xmin) * Cc), float(c * Cc)
        vals = {'mae': s_abs / n_all, 'mae_bb': s_abs_bb / n_bb, 'mae_valid': s_abs / n_valid,
                'mse': s_sq / n_all, 'mse_bb': s_sq_bb / n_bb, 'mse_valid': s_sq / n_valid}
        vals.update(psnr=_psnr(vals['mse']), psnr_bb=_psnr(vals['mse_bb']), psnr_valid=_psnr(vals['mse_valid']))
        if compute_ssim:
            e, g = est[n:n + 1], gt[n:n + 1]
            e_bb, g_bb = e[:, :, ymin:ymax, xmin:xmax], g[:, :, ymin:ymax, xmin:xmax]
            vals['ssim'] = float(_ssim(e, g, data_range=255, size_average=False)[0])
            vals['ssim_bb'] = float(_ssim(e_bb, g_bb, data_range=255, size_average=False)[0])
            # inside the box, pixels outside the mask take the ground-truth value (both are 0 there): identical to ssim_bb on
            # zeroed images -- metric.py:81-84 builds exactly that image
            vals['ssim_valid'] = vals['ssim_bb']
        for k, v in vals.items():
            out[k].append(v)
    res = {}
    for k in _KEYS:
        if out[k]:
            res[k] = np.vstack(out[k])
            res[k + '_mean'] = res[k].mean()
        else:
            res[k] = []
            res[k + '_mean'] = np.nan
    return res
