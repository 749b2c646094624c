// Layout glue between the reference's module-level tensors (NCHW fp32: network.py:251-253,
// train_rnr.py:530-536) and the engine's channels-last 16-bit layouts.  All kernels transpose a
// [C x 32 pixels] tile through shared memory so both the NCHW side (contiguous in w) and the NHWC
// side (contiguous in c) move in full 128-byte lines.
#include "common.cuh"

namespace {

constexpr int TP = 32;   // pixels per tile (along w)

// NCHW fp32 -> fp16 [N,H+2,W+2,Cpad] with reflect halo, channels >= C zero-filled
__global__ void __launch_bounds__(256) pack_nchw_to_act_kernel(const float* __restrict__ src, __half* __restrict__ act,
                                                             __nv_bfloat16* __restrict__ act_b, int N, int C, int Cpad, int H, int W) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float tile[];   // [Cpad][TP+1]
    const int w0 = blockIdx.x * TP, h = blockIdx.y, n = blockIdx.z;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int c = wp; c < Cpad; c += 8) {
        float v = 0.f;
        if (c < C && w0 + lane < W) v = src[(((int64_t)n * C + c) * H + h) * W + w0 + lane];
        tile[c * (TP + 1) + lane] = v;
    }
    __syncthreads();
    const int Hp = H + 2, Wp = W + 2;
    int rows[3], nr = 0;
    rows[nr++] = h + 1;
    if (h == 1) rows[nr++] = 0;
    if (h == H - 2) rows[nr++] = H + 1;
    const int cvecs = Cpad >> 3;
    for (int i = threadIdx.x; i < TP * cvecs; i += 256) {
        const int px = i / cvecs, cv = i % cvecs;
        const int w = w0 + px;
        if (w >= W) continue;
        __align__(16) __half o[8];
        __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float f = tile[(cv * 8 + e) * (TP + 1) + px];
            o[e] = __float2half_rn(f);
            ob[e] = __float2bfloat16_rn(f);
        }
        int cols[3], nc = 0;
        cols[nc++] = w + 1;
        if (w == 1) cols[nc++] = 0;
        if (w == W - 2) cols[nc++] = W + 1;
        for (int a = 0; a < nr; a++)
            for (int b = 0; b < nc; b++) {
                const int64_t o_ = (((int64_t)n * Hp + rows[a]) * Wp + cols[b]) * Cpad + cv * 8;
                *(uint4*)(act + o_) = *(const uint4*)o;
                if (act_b) *(uint4*)(act_b + o_) = *(const uint4*)ob;
            }
    }
}

// fp32 NHWC [N,H,W,ld] -> NCHW [N,C,H,W]
__global__ void __launch_bounds__(256) unpack_nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                int N, int C, int ld, int H, int W) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float tile[];   // [C][TP+1]
    const int w0 = blockIdx.x * TP, h = blockIdx.y, n = blockIdx.z;
    for (int i = threadIdx.x; i < TP * C; i += 256) {
        const int px = i / C, c = i % C;
        if (w0 + px < W) tile[c * (TP + 1) + px] = src[(((int64_t)n * H + h) * W + w0 + px) * ld + c];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int c = wp; c < C; c += 8)
        if (w0 + lane < W) dst[(((int64_t)n * C + c) * H + h) * W + w0 + lane] = tile[c * (TP + 1) + lane];
}

// grad wrt tanh output (NCHW fp32) -> gz = g*(1-t^2) in bf16 zero-halo [N,H+2,W+2,ld]; dbias += sum
// One CTA walks TANH_ROWS image rows of a 32-pixel column strip and keeps its bias partial sums in registers, so the
// per-channel atomics on dbias drop from one per (row, strip) to one per (TANH_ROWS rows, strip).
constexpr int TANH_ROWS = 16;
constexpr int TANH_MAXC8 = 16;       // ld <= 128
__global__ void __launch_bounds__(256) tanh_bwd_pack_kernel(const float* __restrict__ grad, const float* __restrict__ th,
                                                          __nv_bfloat16* __restrict__ gz, float* __restrict__ dbias,
                                                          int N, int C, int ld, int H, int W) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float tile[];   // [ld][TP+1]
    const int w0 = blockIdx.x * TP, n = blockIdx.z;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int Hp = H + 2, Wp = W + 2;
    float bsum[TANH_MAXC8];
#pragma unroll
    for (int j = 0; j < TANH_MAXC8; j++) bsum[j] = 0.f;
    const int h_end = min(H, (int)(blockIdx.y + 1) * TANH_ROWS);
    for (int h = blockIdx.y * TANH_ROWS; h < h_end; h++) {
        for (int c = wp; c < ld; c += 8) {
            float v = 0.f;
            if (c < C && w0 + lane < W) v = grad[(((int64_t)n * C + c) * H + h) * W + w0 + lane];
            tile[c * (TP + 1) + lane] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TP * ld; i += 256) {
            const int px = i / ld, c = i % ld;
            const int w = w0 + px;
            float g = 0.f;
            if (w < W && c < C) {
                const float t = th[(((int64_t)n * H + h) * W + w) * ld + c];
                g = tile[c * (TP + 1) + px] * (1.f - t * t);
            }
            tile[c * (TP + 1) + px] = g;
            if (w < W) gz[(((int64_t)n * Hp + h + 1) * Wp + w + 1) * ld + c] = __float2bfloat16_rn(g);
        }
        __syncthreads();
        if (dbias) {
#pragma unroll
            for (int j = 0; j < TANH_MAXC8; j++) {
                const int c = wp + 8 * j;
                if (c < C) bsum[j] += tile[c * (TP + 1) + lane];
            }
        }
        __syncthreads();
    }
    if (dbias) {
#pragma unroll
        for (int j = 0; j < TANH_MAXC8; j++) {
            const int c = wp + 8 * j;
            if (c < C) {
                float v = bsum[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) atomicAdd(dbias + c, v);
            }
        }
    }
}

// grad w.r.t. reflect-padded input [N,H+2,W+2,ld] -> fold halo -> NCHW fp32 for channels [c0,c0+C)
__global__ void __launch_bounds__(256) fold_to_nchw_kernel(const void* __restrict__ gpad, int dtype, float* __restrict__ dst,
                                                         int N, int C, int c0, int ld, int H, int W,
                                                         const float* __restrict__ add, int nadd) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float tile[];   // [C][TP+1]
    const int w0 = blockIdx.x * TP, h = blockIdx.y, n = blockIdx.z;
    const int Hp = H + 2, Wp = W + 2;
    int rows[3], nr = 0;
    rows[nr++] = h + 1;
    if (h == 1) rows[nr++] = 0;
    if (h == H - 2) rows[nr++] = H + 1;
    for (int i = threadIdx.x; i < TP * C; i += 256) {
        const int px = i / C, c = i % C;
        const int w = w0 + px;
        if (w >= W) continue;
        int cols[3], nc = 0;
        cols[nc++] = w + 1;
        if (w == 1) cols[nc++] = 0;
        if (w == W - 2) cols[nc++] = W + 1;
        float v = 0.f;
        for (int a = 0; a < nr; a++)
            for (int b = 0; b < nc; b++)
                v += ld_any(gpad, (((int64_t)n * Hp + rows[a]) * Wp + cols[b]) * ld + c0 + c, dtype);
        tile[c * (TP + 1) + px] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int c = wp; c < C; c += 8)
        if (w0 + lane < W) {
            float v = tile[c * (TP + 1) + lane];
            if (c < nadd) v += add[(((int64_t)n * nadd + c) * H + h) * W + w0 + lane];     // e.g. the albedo gradient of the tail kernel
            dst[(((int64_t)n * C + c) * H + h) * W + w0 + lane] = v;
        }
}

}  // namespace

extern "C" int rnr_pack_nchw_to_act(const float* src, void* act, void* act_bf16, int N, int C, int Cpad, int H, int W, void* stream) {
    RNR_REQUIRE(Cpad % 8 == 0 && Cpad >= C, "rnr_pack_nchw_to_act: bad Cpad=%d", Cpad);
    RNR_REQUIRE(H >= 2 && W >= 2, "rnr_pack_nchw_to_act: reflect halo needs H,W >= 2");
    const size_t smem = (size_t)Cpad * (TP + 1) * sizeof(float);
    RNR_REQUIRE(smem <= 48 * 1024, "rnr_pack_nchw_to_act: too many channels (%d)", Cpad);
    dim3 grid(rnr_cdiv(W, TP), H, N);
    RNR_PDL_LAUNCH(pack_nchw_to_act_kernel, grid, 256, smem, stream, src, (__half*)act, (__nv_bfloat16*)act_bf16, N, C, Cpad, H, W);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_unpack_nhwc_to_nchw(const float* src, float* dst, int N, int C, int ld, int H, int W, void* stream) {
    const size_t smem = (size_t)C * (TP + 1) * sizeof(float);
    RNR_REQUIRE(smem <= 48 * 1024, "rnr_unpack_nhwc_to_nchw: too many channels (%d)", C);
    dim3 grid(rnr_cdiv(W, TP), H, N);
    RNR_PDL_LAUNCH(unpack_nhwc_to_nchw_kernel, grid, 256, smem, stream, src, dst, N, C, ld, H, W);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_tanh_bwd_pack(const float* grad_nchw, const float* tanh_nhwc, void* gz, float* dbias, int N, int C,
                                 int ld, int H, int W, void* stream) {
    const size_t smem = (size_t)ld * (TP + 1) * sizeof(float);
    RNR_REQUIRE(smem <= 48 * 1024 && ld <= 8 * TANH_MAXC8, "rnr_tanh_bwd_pack: too many channels (%d)", ld);
    dim3 grid(rnr_cdiv(W, TP), rnr_cdiv(H, TANH_ROWS), N);
    RNR_PDL_LAUNCH(tanh_bwd_pack_kernel, grid, 256, smem, stream, grad_nchw, tanh_nhwc, (__nv_bfloat16*)gz, dbias, N, C, ld, H, W);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_fold_to_nchw(const void* gpad, int dtype, float* dst, int N, int C, int c0, int ld, int H, int W,
                                void* stream) {
    const size_t smem = (size_t)C * (TP + 1) * sizeof(float);
    RNR_REQUIRE(smem <= 48 * 1024, "rnr_fold_to_nchw: too many channels (%d)", C);
    dim3 grid(rnr_cdiv(W, TP), H, N);
    RNR_PDL_LAUNCH(fold_to_nchw_kernel, grid, 256, smem, stream, gpad, dtype, dst, N, C, c0, ld, H, W, nullptr, 0);
    RNR_LAUNCH_CHECK();
    return 0;
}

// same, plus dst[:, :nadd] += add ([N, nadd, H, W] fp32): folds `gi[:, :6] += g_alb` of the fused step into this pass
extern "C" int rnr_fold_to_nchw_add(const void* gpad, int dtype, float* dst, int N, int C, int c0, int ld, int H, int W,
                                    const float* add, int nadd, void* stream) {
    const size_t smem = (size_t)C * (TP + 1) * sizeof(float);
    RNR_REQUIRE(smem <= 48 * 1024, "rnr_fold_to_nchw_add: too many channels (%d)", C);
    RNR_REQUIRE(nadd >= 0 && nadd <= C && (nadd == 0 || add), "rnr_fold_to_nchw_add: bad add tensor");
    dim3 grid(rnr_cdiv(W, TP), H, N);
    RNR_PDL_LAUNCH(fold_to_nchw_kernel, grid, 256, smem, stream, gpad, dtype, dst, N, C, c0, ld, H, W, add, nadd);
    RNR_LAUNCH_CHECK();
    return 0;
}
