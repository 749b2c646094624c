#!/usr/bin/env python
"""Build the reference's OWN CUDA extension (neural_renderer.cuda.{rasterize, load_textures, create_texture_image}) for sm_100a
from a patched BUILD COPY under the git-ignored baseline/_ref/nr_ext/ (the torch-1.1 sources need four mechanical
host-glue substitutions to compile against torch 2.x; the __global__ kernel bodies are untouched):

    .type(), "name"  -> .scalar_type(), "name"      (AT_DISPATCH_FLOATING_TYPES)
    .data<T>()        -> .data_ptr<T>()
    AT_CHECK          -> TORCH_CHECK, x.type().is_cuda() -> x.is_cuda()
    <torch/torch.h>   -> <torch/extension.h>

It is TEST / BENCH infrastructure: the GPU-side oracle of the seven B2 entry points (tests/test_b2_gpu.py) and "the kernel
our rasterizer must beat" of bench.py --config rnr_infer.  Nothing is copied into tracked paths; the product never loads it.
Runs in the build container (needs /root/reference; nvcc cross-compiles without a GPU); the built .so files travel to the GPU box.
"""
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/neural_renderer/neural_renderer/cuda'
DST = os.path.join(ROOT, 'baseline', '_ref', 'nr_ext')
EXTS = {
    'ref_rasterize': ['rasterize_cuda.cpp', 'rasterize_cuda_kernel.cu'],
    'ref_load_textures': ['load_textures_cuda.cpp', 'load_textures_cuda_kernel.cu'],
    'ref_create_texture_image': ['create_texture_image_cuda.cpp', 'create_texture_image_cuda_kernel.cu'],
}


def patch(txt):
    txt = re.sub(r'\.type\(\)\s*,\s*"', '.scalar_type(), "', txt)
    txt = re.sub(r'\.data<', '.data_ptr<', txt)
    txt = txt.replace('.type().is_cuda()', '.is_cuda()')
    txt = txt.replace('AT_CHECK', 'TORCH_CHECK')
    txt = txt.replace('<torch/torch.h>', '<torch/extension.h>')
    return txt


def so_path(name):
    return os.path.join(DST, name, name + '.so')


def build(verbose=True):
    if not os.path.isdir(SRC):
        ok = all(os.path.exists(so_path(n)) for n in EXTS)
        if verbose:
            print('build_ref_ext: %s absent; prebuilt extension %s' % (SRC, 'present' if ok else 'MISSING'))
        return ok
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils.cpp_extension import load
    for name, files in EXTS.items():
        bdir = os.path.join(DST, name)
        os.makedirs(bdir, exist_ok=True)
        srcs = []
        for f in files:
            d = os.path.join(bdir, f)
            new = patch(open(os.path.join(SRC, f)).read())
            if not os.path.exists(d) or open(d).read() != new:
                open(d, 'w').write(new)
            srcs.append(d)
        if os.path.exists(so_path(name)) and all(os.path.getmtime(so_path(name)) >= os.path.getmtime(s) for s in srcs):
            continue
        load(name=name, sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False,
             extra_cuda_cflags=['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo'])
        if verbose:
            print('built', so_path(name))
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
