"""``neural_renderer.cuda.load_textures`` (load_textures_cuda.cpp:20-39): cold, set-up only, outside the hot path."""


def load_textures(*a, **k):
    raise NotImplementedError('neural_renderer.cuda.load_textures is outside the relighting hot path (load_obj is always called '
                              'with load_texture=False: network.py:106); librnr_b200 does not provide it')
