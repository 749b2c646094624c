"""Drop-in for the reference's ``misc`` module (misc.py:5-73) on librnr_b200.so."""
import numpy as np
import torch

from .. import ops
from ._dev import on_cuda


def interpolate_bilinear(data, sub_x, sub_y):
    """data [H,W,C]; sub_x, sub_y [...] pixel coordinates -> [..., C]   (misc.py:5-42).

    4-tap gather with the reference's hard validity mask (any coordinate outside [0,W-1]x[0,H-1] gives
    exactly 0) and its right/bottom edge weight fix-up.  Differentiable w.r.t. ``data``.  Tensors that
    live on the host (the reference scripts call this at set-up time with CPU tensors, e.g.
    train_rnr.py:307-308) are staged through the GPU: the arithmetic always runs in the CUDA kernel."""
    (data, sub_x, sub_y), back = on_cuda(data, sub_x, sub_y)
    return back(ops.interpolate_bilinear(data, sub_x, sub_y))


def interpolate_bilinear_np(data, sub_x, sub_y):
    """numpy flavour (misc.py:45-73): no validity mask, no edge fix-up -- plain clamped 4-tap blend.
    Host-side helper of the reference's offline tools; evaluated with the same CUDA gather kernel on
    clamped coordinates is not equivalent at the borders, so this one is plain numpy index arithmetic."""
    x0 = np.floor(sub_x).astype(np.int64)
    y0 = np.floor(sub_y).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    H, W = data.shape[0], data.shape[1]
    x0, x1 = np.clip(x0, 0, W - 1), np.clip(x1, 0, W - 1)
    y0, y1 = np.clip(y0, 0, H - 1), np.clip(y1, 0, H - 1)
    ax, bx, ay, by = x1 - sub_x, sub_x - x0, y1 - sub_y, sub_y - y0
    return (data[y0, x0] * (ax * ay)[..., None] + data[y1, x0] * (ax * by)[..., None] +
            data[y0, x1] * (bx * ay)[..., None] + data[y1, x1] * (bx * by)[..., None])
