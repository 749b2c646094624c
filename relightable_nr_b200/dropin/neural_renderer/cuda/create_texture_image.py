"""``neural_renderer.cuda.create_texture_image`` (create_texture_image_cuda.cpp:18-33): tiled atlas image from per-face texture cubes.
Cold (save_obj only; never called by the train / test scripts); csrc/nr_cold.cu."""
import ctypes as C

import torch

from .... import _lib

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
_lib.register_sigs({"rnr_nr_create_texture_image": [vp, vp, vp, i64, i32, i32, i32, i32, f32, vp]})


def create_texture_image(vertices_all, textures, image, eps):
    """vertices_all [nf,3,2] f32 (pixel corners of each face's tile), textures [nf,ts,ts,ts,3] f32, image [h,w,3] f32 (filled in
    place: tile width = int(sqrt(nf - 1)) + 1, tile size = w / tile width) -> image."""
    for t, name in ((vertices_all, 'vertices_all'), (textures, 'textures'), (image, 'image')):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError('%s must be a contiguous CUDA float32 tensor' % name)
    nf = int(textures.shape[0])
    tile_width = int((nf - 1) ** 0.5) + 1
    tso = int(image.shape[1]) // tile_width
    _lib.check(_lib.lib().rnr_nr_create_texture_image(vertices_all.data_ptr(), textures.data_ptr(), image.data_ptr(), image.numel(), nf,
                                                      int(textures.shape[1]), tso, tile_width, float(eps),
                                                      torch.cuda.current_stream().cuda_stream), 'rnr_nr_create_texture_image')
    return image
