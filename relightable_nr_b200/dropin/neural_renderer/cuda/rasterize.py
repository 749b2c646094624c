"""``neural_renderer.cuda.rasterize``: same five entry points as cuda/rasterize_cuda.cpp:70-199, caller-allocated
contiguous CUDA tensors filled in place and returned.  ``forward_face_index_map`` runs the tile-culled B200 kernels;
the rgb sampling and the three backward functions are outside the hot path (the reference never consumes the rgb of
its all-zero face texture, network.py:157, and never differentiates the rasterizer) and raise instead of silently
returning something else."""
import ctypes as C

import torch

from .... import _lib

vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
_lib.register_sigs({
    "rnr_raster_face_setup": [vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, i32, vp],
    "rnr_raster_tiles": [vp, vp, vp, i32, i32, f32, f32, i32, vp, vp, vp, vp, vp, vp, i32, vp],
})


def _check(t, name, dtype=torch.float32):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError('%s must be a CUDA tensor' % name)          # CHECK_CUDA (rasterize_cuda.cpp:66)
    if not t.is_contiguous():
        raise RuntimeError('%s must be contiguous' % name)             # CHECK_CONTIGUOUS (:67)
    if t.dtype != dtype:
        raise RuntimeError('%s must be %s' % (name, dtype))


def forward_face_index_map(faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv, image_size, near, far,
                           return_rgb, return_alpha, return_depth):
    """rasterize_cuda.cpp:70-95.  Every pixel of the three maps is written (background: -1 / 0 / far)."""
    _check(faces, 'faces'); _check(face_index_map, 'face_index_map', torch.int32); _check(weight_map, 'weight_map')
    _check(depth_map, 'depth_map'); _check(face_inv_map, 'face_inv_map'); _check(faces_inv, 'faces_inv')
    N, nf = faces.shape[0], faces.shape[1]
    s = torch.cuda.current_stream().cuda_stream
    bbox = torch.empty((N, nf, 2), dtype=torch.int32, device=faces.device)
    L = _lib.lib()
    _lib.check(L.rnr_raster_face_setup(None, 0, None, 1, faces.data_ptr(), nf, int(image_size), None, faces_inv.data_ptr(),
                                       bbox.data_ptr(), N, s), 'rnr_raster_face_setup')
    want_inv = bool(return_depth) and face_inv_map.numel() == N * int(image_size) ** 2 * 9
    _lib.check(L.rnr_raster_tiles(faces.data_ptr(), faces_inv.data_ptr(), bbox.data_ptr(), nf, int(image_size), float(near),
                                  float(far), 0, face_index_map.data_ptr(), weight_map.data_ptr(), depth_map.data_ptr(), None,
                                  face_inv_map.data_ptr() if want_inv else None, None, N, s), 'rnr_raster_tiles')
    return [face_index_map, weight_map, depth_map, face_inv_map]


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(
            'neural_renderer.cuda.rasterize.%s is outside the relighting hot path (the rasterizer is forward-only and its rgb '
            'output is never consumed: SURVEY.md 8a rows a3/a4); librnr_b200 does not provide it' % name)
    f.__name__ = name
    return f


forward_texture_sampling = _out_of_scope('forward_texture_sampling')
backward_pixel_map = _out_of_scope('backward_pixel_map')
backward_textures = _out_of_scope('backward_textures')
backward_depth_map = _out_of_scope('backward_depth_map')
