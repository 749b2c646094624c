"""``neural_renderer.cuda.rasterize``: same five entry points as cuda/rasterize_cuda.cpp:70-199, caller-allocated
contiguous CUDA tensors filled in place and returned.  ``forward_face_index_map`` runs the tile-culled B200 kernels;
the rgb sampling and the three backward functions are outside the hot path (the reference never consumes the rgb of
its all-zero face texture, network.py:157, and never differentiates the rasterizer); they run the kernels of csrc/nr_cold.cu,
which restate the reference's and are checked against it on the GPU (tests/test_b2_gpu.py)."""
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


_lib.register_sigs({
    "rnr_nr_forward_texture_sampling": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp],
    "rnr_nr_backward_pixel_map": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, i32, vp],
    "rnr_nr_backward_textures": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "rnr_nr_backward_depth_map": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
})


def _stream():
    return torch.cuda.current_stream().cuda_stream


def forward_texture_sampling(faces, textures, face_index_map, weight_map, depth_map, rgb_map, sampling_index_map, sampling_weight_map,
                             image_size, eps):
    """rasterize_cuda.cpp:97-122: per-pixel trilinear sample of the face's texture cube [B,nf,ts,ts,ts,3]; fills rgb_map and the 8
    sampling indices / weights per pixel in place.  Cold: the relighting path discards this output (network.py:157)."""
    _check(faces, 'faces'); _check(textures, 'textures'); _check(face_index_map, 'face_index_map', torch.int32)
    _check(weight_map, 'weight_map'); _check(depth_map, 'depth_map'); _check(rgb_map, 'rgb_map')
    _check(sampling_index_map, 'sampling_index_map', torch.int32); _check(sampling_weight_map, 'sampling_weight_map')
    B, nf = faces.shape[0], faces.shape[1]
    _lib.check(_lib.lib().rnr_nr_forward_texture_sampling(
        faces.data_ptr(), textures.data_ptr(), face_index_map.data_ptr(), weight_map.data_ptr(), depth_map.data_ptr(), rgb_map.data_ptr(),
        sampling_index_map.data_ptr(), sampling_weight_map.data_ptr(), B, nf, int(image_size), int(textures.shape[2]), float(eps), _stream()),
        'rnr_nr_forward_texture_sampling')
    return [rgb_map, sampling_index_map, sampling_weight_map]


def backward_pixel_map(faces, face_index_map, rgb_map, alpha_map, grad_rgb_map, grad_alpha_map, grad_faces, image_size, eps, return_rgb,
                       return_alpha):
    """rasterize_cuda.cpp:124-148: silhouette gradient w.r.t. the projected faces (grad_faces [B,nf,3,3], written in place)."""
    _check(faces, 'faces'); _check(face_index_map, 'face_index_map', torch.int32); _check(rgb_map, 'rgb_map'); _check(alpha_map, 'alpha_map')
    _check(grad_rgb_map, 'grad_rgb_map'); _check(grad_alpha_map, 'grad_alpha_map'); _check(grad_faces, 'grad_faces')
    B, nf = faces.shape[0], faces.shape[1]
    _lib.check(_lib.lib().rnr_nr_backward_pixel_map(
        faces.data_ptr(), face_index_map.data_ptr(), rgb_map.data_ptr(), alpha_map.data_ptr(), grad_rgb_map.data_ptr(), grad_alpha_map.data_ptr(),
        grad_faces.data_ptr(), B, nf, int(image_size), float(eps), int(return_rgb), int(return_alpha), _stream()), 'rnr_nr_backward_pixel_map')
    return grad_faces


def backward_textures(face_index_map, sampling_weight_map, sampling_index_map, grad_rgb_map, grad_textures, num_faces):
    """rasterize_cuda.cpp:150-167: grad_textures [B,nf,ts,ts,ts,3] += scatter of grad_rgb_map through the sampling maps."""
    _check(face_index_map, 'face_index_map', torch.int32); _check(sampling_weight_map, 'sampling_weight_map')
    _check(sampling_index_map, 'sampling_index_map', torch.int32); _check(grad_rgb_map, 'grad_rgb_map'); _check(grad_textures, 'grad_textures')
    B, is_ = face_index_map.shape[0], face_index_map.shape[1]
    _lib.check(_lib.lib().rnr_nr_backward_textures(
        face_index_map.data_ptr(), sampling_weight_map.data_ptr(), sampling_index_map.data_ptr(), grad_rgb_map.data_ptr(), grad_textures.data_ptr(),
        B, int(num_faces), is_, int(grad_textures.shape[2]), _stream()), 'rnr_nr_backward_textures')
    return grad_textures


def backward_depth_map(faces, depth_map, face_index_map, face_inv_map, weight_map, grad_depth_map, grad_faces, image_size):
    """rasterize_cuda.cpp:169-191: grad_faces += d depth / d projected vertices."""
    _check(faces, 'faces'); _check(depth_map, 'depth_map'); _check(face_index_map, 'face_index_map', torch.int32)
    _check(face_inv_map, 'face_inv_map'); _check(weight_map, 'weight_map'); _check(grad_depth_map, 'grad_depth_map'); _check(grad_faces, 'grad_faces')
    B, nf = faces.shape[0], faces.shape[1]
    _lib.check(_lib.lib().rnr_nr_backward_depth_map(
        faces.data_ptr(), depth_map.data_ptr(), face_index_map.data_ptr(), face_inv_map.data_ptr(), weight_map.data_ptr(), grad_depth_map.data_ptr(),
        grad_faces.data_ptr(), B, nf, int(image_size), _stream()), 'rnr_nr_backward_depth_map')
    return grad_faces
