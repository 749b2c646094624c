"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libref_raster.so: the reference's own rasterizer kernels
(neural_renderer/cuda/rasterize_cuda_kernel.cu:24-169) compiled for the CPU by oracle/build_oracle.py.
``forward_face_index_map`` mirrors RasterizeFunction.forward_face_index_map (rasterize.py:161-168) including the
pre-filled output buffers of rasterize.py:50-69."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libref_raster.so')
_lib = None


def available():
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.ref_forward_face_index_map.argtypes = [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_float] * 2 + [C.c_int]
        _lib.ref_forward_face_index_map.restype = None
    return _lib


def forward_face_index_map(faces, image_size, near=0.0, far=1e5, return_depth=True):
    """faces [B,nf,3,3] float32 -> (face_index_map [B,is,is] i32, weight_map [B,is,is,3], depth_map [B,is,is],
    face_inv_map [B,is,is,9], faces_inv [B,nf,9]); unflipped."""
    f = np.ascontiguousarray(faces, dtype=np.float32)
    B, nf = f.shape[:2]
    is_ = int(image_size)
    faces_inv = np.zeros((B, nf, 9), dtype=np.float32)
    fim = np.full((B, is_, is_), -1, dtype=np.int32)
    wm = np.zeros((B, is_, is_, 3), dtype=np.float32)
    dm = np.full((B, is_, is_), far, dtype=np.float32)
    fiv = np.zeros((B, is_, is_, 9), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _load().ref_forward_face_index_map(p(f), p(faces_inv), p(fim), p(wm), p(dm), p(fiv), B, nf, is_, near, far, int(return_depth))
    return fim, wm, dm, fiv, faces_inv
