// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16, SS operands, SWIZZLE_128B K-major) for N = 64/128/256,
// same accumulator vs alternating accumulators, SBO 1024 vs 1280.  No loads: operands are whatever is in shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench.bin tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../relightable_nr_b200/csrc/tc_ptx.cuh"
void rnr_set_error(const char*, ...) {}
void rnr_count_launch(void) {}

__global__ void __launch_bounds__(128, 1) bench(int n, int reps, int alt, int sbo_a, int kadv, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(128, n, RNR_F16, RNR_F16, 0, 0);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        long long t0 = 0, t1 = 0;
        for (int round = 0; round < 2; round++) {
            t0 = clock64();
            if (elect_one_sync()) {
                for (int i = 0; i < reps; i++) {
                    const uint32_t off = kadv ? (uint32_t)((i & 3) * 2) : 0u;
                    uint64_t da = 0;
                    da |= (uint64_t)(((a_base + ((i >> 2) % 9) * 128u) & 0x3FFFF) >> 4);
                    da |= (uint64_t)1 << 16; da |= (uint64_t)((uint32_t)sbo_a >> 4) << 32; da |= (uint64_t)1 << 46; da |= (uint64_t)2 << 61;
                    const uint64_t db = make_kmajor_desc(b_base + ((i >> 2) % 3) * (uint32_t)n * 128u, 64);
                    umma_f16(tm + (alt ? (uint32_t)(i & 1) * 256u : 0u), da + off, db + off, idesc, 1);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, round & 1);
            t1 = clock64();
        }
        if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 512;
    for (int grid : {1, 148})
        for (int n : {64, 128, 256})
            for (int alt : {0, 1})
                for (int sbo : {1024, 1280})
                    for (int kadv : {0, 1}) {
                        bench<<<grid, 128, 200 * 1024>>>(n, reps, alt, sbo, kadv, d);
                        long long h = 0;
                        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                        cudaError_t e = cudaGetLastError();
                        printf("grid %3d N %3d alt %d sbo %4d kadv %d : %6.1f cycles/MMA (ideal %d)%s\n", grid, n, alt, sbo, kadv, (double)h / reps, n / 2,
                               e == cudaSuccess ? "" : cudaGetErrorString(e));
                    }
    return 0;
}
