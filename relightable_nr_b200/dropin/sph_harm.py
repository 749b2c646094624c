"""Drop-in for the reference's ``sph_harm`` module (sph_harm.py:6-102) on librnr_b200.so.

``evaluate_sh_basis`` replaces the pyshtools (CPU, Fortran) evaluation with a CUDA kernel: real,
4pi-orthonormal basis, no Condon-Shortley phase, order (l, m=-l..l), m<0 <-> sin(|m| phi) -- the
convention of ``SHCoeffs.from_zeros(l, csphase=1, normalization='ortho')`` (sph_harm.py:66-68)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib, ops
from ._dev import on_cuda

_lib.register_sigs({"rnr_sh_basis": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]})


def cart2sph(x, y, z):
    t = torch if isinstance(x, torch.Tensor) else np
    atan2 = torch.atan2 if t is torch else np.arctan2
    rxy2 = x ** 2 + y ** 2
    return atan2(y, x), atan2(z, t.sqrt(rxy2)), t.sqrt(rxy2 + z ** 2)


def sph2cart(azimuth, elevation, r):
    fa = torch if isinstance(azimuth, torch.Tensor) else np
    fe = torch if isinstance(elevation, torch.Tensor) else np
    ce = fe.cos(elevation)
    return r * ce * fa.cos(azimuth), r * ce * fa.sin(azimuth), r * fe.sin(elevation)


def evaluate_sh_basis(lmax=0, azi=None, pol=None, directions=None):
    """-> np.ndarray [num_sample, (lmax+1)^2] float64.  ``directions`` [num_sample,3] (numpy or tensor), or
    ``azi``/``pol`` in degrees (sph_harm.py:41-71)."""
    if azi is not None or pol is not None:
        a, p = np.deg2rad(np.asarray(azi, dtype=np.float64)), np.deg2rad(np.asarray(pol, dtype=np.float64))
        directions = np.stack((np.sin(p) * np.cos(a), np.sin(p) * np.sin(a), np.cos(p)), -1)
    if isinstance(directions, torch.Tensor):
        d = directions.detach()
    else:
        d = torch.from_numpy(np.ascontiguousarray(np.asarray(directions, dtype=np.float32)))
    (d,), _ = on_cuda(d.float())
    d = d.contiguous()
    n = d.shape[0]
    out = torch.empty((n, (lmax + 1) ** 2), dtype=torch.float64, device=d.device)
    _lib.check(_lib.lib().rnr_sh_basis(d.data_ptr(), out.data_ptr(), n, int(lmax), torch.cuda.current_stream().cuda_stream),
               'rnr_sh_basis')
    return out.cpu().numpy()


def evaluate_sh_basis_l2(directions):
    """[..., 3] CUDA tensor -> [..., 9] fp32 on device (the per-pixel map of test_rnr.py:322-329 without the
    GPU->CPU->GPU round trip)."""
    return ops.sh_basis_l2(directions)


def _as_tensor(a):
    return (a, False) if isinstance(a, torch.Tensor) else (torch.from_numpy(np.asarray(a, dtype=np.float32)), True)


def fit_sh_coeff(samples, sh_basis_val):
    """[num_sample,C] or [L,num_sample,C] x [num_sample,num_basis] -> [num_basis,C] or [L,num_basis,C] (sph_harm.py:74-88)."""
    s, was_np = _as_tensor(samples)
    b, _ = _as_tensor(sh_basis_val)
    (s, b), back = on_cuda(s, b)
    out = back(ops.sh_fit(s, b))
    return out.numpy() if was_np else out


def reconstruct_sh(sh_coeff, sh_basis_val):
    """[num_basis,C] or [L,num_basis,C] x [num_sample,num_basis] -> [num_sample,C] or [L,num_sample,C] (sph_harm.py:91-102)."""
    c, was_np = _as_tensor(sh_coeff)
    b, _ = _as_tensor(sh_basis_val)
    (c, b), back = on_cuda(c, b)
    out = back(ops.sh_reconstruct(c, b))
    return out.numpy() if was_np else out
