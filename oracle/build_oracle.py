"""TEST INFRASTRUCTURE ONLY -- builds the checker binaries under oracle/_ref/ (git-ignored, shipped to the GPU box).

libref_raster.so: the reference's OWN rasterizer kernel bodies
    /root/reference/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu   (kernel 1 :24-68, kernel 2 :70-169)
compiled for the host CPU.  The two ``__global__`` templates are plain C++ apart from blockIdx/blockDim/threadIdx, so a
shim translation unit defines those as thread-local variables, includes the kernels' text (cut out of the source file where
it lies, into a temporary directory -- nothing is copied into this repository) and drives them over the launch grid with
OpenMP.  The ATen host wrappers of that file (which need torch 1.x) are not compiled.
Floating-point contraction is disabled (-ffp-contract=off): the result is the kernels' arithmetic as written.

Runs only where /root/reference exists (the build container); elsewhere the prebuilt file is used as is.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')
REF_CU = '/root/reference/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu'

SHIM_HEAD = r'''
#include <cstdint>
#include <cmath>
#include <cstring>
struct idx3 { int x, y, z; };
static thread_local idx3 blockIdx, blockDim, threadIdx;
#define __global__
#define __device__
// CUDA's mixed float/double min/max overloads promote to double
template <class A, class B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }
template <class A, class B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
namespace refk {
'''

SHIM_TAIL = r'''
}  // namespace refk

extern "C" void ref_forward_face_index_map(const float* faces, float* faces_inv, int32_t* face_index_map, float* weight_map,
                                           float* depth_map, float* face_inv_map, int batch_size, int num_faces, int image_size,
                                           float near, float far, int return_depth) {
    const int threads = 512;
    const long n1 = (long)batch_size * num_faces;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n1; i++) {
        blockDim = {threads, 1, 1};
        blockIdx = {(int)(i / threads), 0, 0};
        threadIdx = {(int)(i % threads), 0, 0};
        refk::forward_face_index_map_cuda_kernel_1<float>(faces, faces_inv, batch_size, num_faces, image_size);
    }
    const long n2 = (long)batch_size * image_size * image_size;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < n2; i++) {
        blockDim = {threads, 1, 1};
        blockIdx = {(int)(i / threads), 0, 0};
        threadIdx = {(int)(i % threads), 0, 0};
        refk::forward_face_index_map_cuda_kernel_2<float>(faces, faces_inv, face_index_map, weight_map, depth_map, face_inv_map,
                                                         batch_size, num_faces, image_size, near, far, 0, 1, return_depth);
    }
}
'''


def _kernel_text():
    """Text of the first two ``template <typename scalar_t> __global__`` kernels of the reference file."""
    lines = open(REF_CU).read().split('\n')
    starts = [i for i, ln in enumerate(lines) if ln.strip() == 'template <typename scalar_t>']
    assert len(starts) >= 3, 'unexpected layout of %s' % REF_CU
    body = '\n'.join(lines[starts[0]:starts[2]])
    assert 'forward_face_index_map_cuda_kernel_1' in body and 'forward_face_index_map_cuda_kernel_2' in body
    assert 'forward_texture_sampling' not in body
    return body


def build(verbose=True):
    out = os.path.join(OUT_DIR, 'libref_raster.so')
    if not os.path.exists(REF_CU):
        if verbose:
            print('oracle/_ref: /root/reference absent -- using prebuilt checker binaries' if os.path.exists(out)
                  else 'oracle/_ref: /root/reference absent and no prebuilt libref_raster.so (reference-kernel tests will skip)')
        return out if os.path.exists(out) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(REF_CU), os.path.getmtime(__file__)):
        return out
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'ref_raster_cpu.cpp')
        with open(src, 'w') as fh:
            fh.write(SHIM_HEAD + _kernel_text() + SHIM_TAIL)
        cmd = ['g++', '-O2', '-std=c++14', '-shared', '-fPIC', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', src, '-o', out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('building the reference rasterizer kernels for the CPU failed:\n' + r.stderr)
    if verbose:
        print('built', out)
    return out


if __name__ == '__main__':
    build(verbose=True)
