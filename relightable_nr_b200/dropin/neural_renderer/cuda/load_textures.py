"""``neural_renderer.cuda.load_textures`` (load_textures_cuda.cpp:20-39): per-face texture cubes from a material image.  Cold, set-up
only (load_obj is always called with load_texture=False on the relighting path: network.py:106); csrc/nr_cold.cu."""
import ctypes as C

import torch

from .... import _lib

vp, i32 = C.c_void_p, C.c_int
_lib.register_sigs({"rnr_nr_load_textures": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]})


def load_textures(image, faces, textures, is_update, texture_wrapping, use_bilinear):
    """image [H,W,3] f32, faces [nf,3,2] f32 (uv, wrapped in place), textures [nf,ts,ts,ts,3] f32 (filled in place where
    is_update [nf] i32 != 0), texture_wrapping 0..3 (REPEAT, MIRRORED_REPEAT, CLAMP_TO_EDGE, CLAMP_TO_BORDER) -> textures."""
    for t, name, dt in ((image, 'image', torch.float32), (faces, 'faces', torch.float32), (textures, 'textures', torch.float32),
                        (is_update, 'is_update', torch.int32)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dt):
            raise RuntimeError('%s must be a contiguous CUDA %s tensor' % (name, dt))
    _lib.check(_lib.lib().rnr_nr_load_textures(image.data_ptr(), faces.data_ptr(), textures.data_ptr(), is_update.data_ptr(),
                                               int(textures.shape[0]), int(textures.shape[1]), int(image.shape[0]), int(image.shape[1]),
                                               int(texture_wrapping), int(bool(use_bilinear)), torch.cuda.current_stream().cuda_stream),
               'rnr_nr_load_textures')
    return textures
