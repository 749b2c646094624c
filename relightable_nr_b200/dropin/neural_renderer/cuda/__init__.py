"""``neural_renderer.cuda``: the three extension modules of the reference (rasterize, load_textures,
create_texture_image; cuda/rasterize_cuda.cpp:193-199, load_textures_cuda.cpp:37-39, create_texture_image_cuda.cpp:31-33)."""
from . import rasterize, load_textures, create_texture_image  # noqa: F401
