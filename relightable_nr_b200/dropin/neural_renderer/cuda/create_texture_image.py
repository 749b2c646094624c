"""``neural_renderer.cuda.create_texture_image`` (create_texture_image_cuda.cpp:18-33): used by save_obj only, outside the hot path."""


def create_texture_image(*a, **k):
    raise NotImplementedError('neural_renderer.cuda.create_texture_image is outside the relighting hot path (save_obj is never '
                              'called by the train/test scripts); librnr_b200 does not provide it')
