// BatchNorm2d with *batch statistics* (the only mode the reference ever runs: SURVEY 3.3,
// test_rnr.py:229-233, train_rnr.py:398-405) + LeakyReLU/ReLU + Dropout2d, forward and backward.
// Reference operators: pytorch_prototyping/pytorch_prototyping.py:177-197 (UpBlock),
// :250-272 (DownBlock), :470-476 (Unet.in_layer).
//
// HBM-bound elementwise/reduction kernels: 128-bit loads, 8 channels per thread, channels-last.
#include "common.cuh"
#include <stdlib.h>

namespace {

// -------------------------------------------------------------------------------------------
// forward statistics finalize
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) bn_finalize_kernel(const float* __restrict__ partials, int T, int ld, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float* mean, float* invstd, float* scale, float* shift,
                                   float* running_mean, float* running_var, float momentum) {
    __shared__ double s_s[16][33], s_q[16][33];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;   // 16 warps
    const int c = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int t = w; t < T; t += 16) {
            s += (double)partials[((int64_t)t * 2 + 0) * ld + c];
            q += (double)partials[((int64_t)t * 2 + 1) * ld + c];
        }
    }
    s_s[w][lane] = s; s_q[w][lane] = q;
    __syncthreads();
    if (w == 0 && c < C) {
        for (int i = 1; i < 16; i++) { s += s_s[i][lane]; q += s_q[i][lane]; }
        const double m = s / count;
        double var = q / count - m * m;
        if (var < 0.0) var = 0.0;
        const float istd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        mean[c] = (float)m;
        invstd[c] = istd;
        scale[c] = g * istd;
        shift[c] = b - (float)m * g * istd;
        if (running_mean) {
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
}

// -------------------------------------------------------------------------------------------
// forward apply: act = drop * lrelu(raw*scale + shift), fp16, reflect halo
// -------------------------------------------------------------------------------------------
// 8 consecutive channels of the raw (pre-BatchNorm) convolution output: fp32 (2 x 16 bytes) or 16-bit (16 bytes)
__device__ __forceinline__ void load_raw8(const void* raw, int64_t idx, int dtype, float* r) {
    if (dtype == RNR_F32) {
        const float4 a = __ldcs((const float4*)((const float*)raw + idx));
        const float4 b = __ldcs((const float4*)((const float*)raw + idx + 4));
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    } else {
        const uint4 u = __ldcs((const uint4*)((const unsigned short*)raw + idx));
        const unsigned short* us = (const unsigned short*)&u;
#pragma unroll
        for (int e = 0; e < 8; e++) r[e] = cvt16(us[e], dtype);
    }
}

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const void* __restrict__ raw, int raw_dtype, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ drop, float slope,
                                  __half* __restrict__ act, __nv_bfloat16* __restrict__ act_b, int N, int H, int W, int C) {
    pdl_launch_dependents();
    pdl_wait();
    // grid.y = image row (n*H + h); threads of a row = (w, 8-channel vector): 32-bit index math only
    const unsigned vpp = (unsigned)C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)W * vpp) return;
    const int row = blockIdx.y;
    const int n = row / H, h = row - n * H;
    const int w = (int)(idx / vpp);
    const int c = (int)(idx - (unsigned)w * vpp) * 8;
    const int Hp = H + 2, Wp = W + 2;
    const int64_t pix = (int64_t)row * W + w;
    float rr[8];
    load_raw8(raw, pix * C + c, raw_dtype, rr);
    const float4 s0 = *(const float4*)(scale + c), s1 = *(const float4*)(scale + c + 4);
    const float4 t0 = *(const float4*)(shift + c), t1 = *(const float4*)(shift + c + 4);
    float v[8] = {rr[0] * s0.x + t0.x, rr[1] * s0.y + t0.y, rr[2] * s0.z + t0.z, rr[3] * s0.w + t0.w,
                  rr[4] * s1.x + t1.x, rr[5] * s1.y + t1.y, rr[6] * s1.z + t1.z, rr[7] * s1.w + t1.w};
    float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (drop) {
        const float4 d0 = *(const float4*)(drop + n * C + c), d1 = *(const float4*)(drop + n * C + c + 4);
        dm[0] = d0.x; dm[1] = d0.y; dm[2] = d0.z; dm[3] = d0.w; dm[4] = d1.x; dm[5] = d1.y; dm[6] = d1.z; dm[7] = d1.w;
    }
    __align__(16) __half o[8];
    __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        float z = v[e];
        z = z > 0.f ? z : z * slope;
        z *= dm[e];
        o[e] = __float2half_rn(z);
        ob[e] = __float2bfloat16_rn(z);
    }
    const uint4 ov = *(const uint4*)o;
    const uint4 ovb = *(const uint4*)ob;
    // rows / cols this pixel lands on in the padded tensor (reflect halo)
    int rows[3], cols[3], nr = 0, nc = 0;
    rows[nr++] = h + 1;
    if (h == 1) rows[nr++] = 0;
    if (h == H - 2) rows[nr++] = H + 1;
    cols[nc++] = w + 1;
    if (w == 1) cols[nc++] = 0;
    if (w == W - 2) cols[nc++] = W + 1;
    for (int a = 0; a < nr; a++)
        for (int b = 0; b < nc; b++) {
            const int64_t o_ = (((int64_t)n * Hp + rows[a]) * Wp + cols[b]) * C + c;
            *(uint4*)(act + o_) = ov;
            if (act_b) *(uint4*)(act_b + o_) = ovb;
        }
}

// The same pass fed by fp64 TOTALS of the batch sums (rnr_conv_plan_set_stat_totals) instead of finalized scale / shift arrays:
// every block turns the totals into the [scale | shift] table of the layer in shared memory (one channel per thread: mean, biased
// variance, invstd in double -- the arithmetic of bn_finalize_kernel), block (0,0) also publishes mean / invstd / scale / shift for
// the backward pass and updates the running statistics, and the last block to have read the totals (ticket) re-zeroes them.
// A block covers RB image rows, so that a 512^2 layer runs ~2000 blocks, not 8192.  No finalize launch between the convolution
// and this kernel.
__global__ void __launch_bounds__(256) bn_act_fwd_tot_kernel(const void* __restrict__ raw, int raw_dtype, double* __restrict__ totals,
                                  int* __restrict__ ticket, double count, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                  float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ running_mean,
                                  float* __restrict__ running_var, float momentum, const float* __restrict__ drop, float slope,
                                  __half* __restrict__ act, __nv_bfloat16* __restrict__ act_b, int N, int H, int W, int C, int RB) {
    extern __shared__ __align__(16) float s_tab[];      // [2][C]: scale, shift
    __shared__ int s_last;
    pdl_launch_dependents();
    pdl_wait();
    const bool first = (blockIdx.x == 0 && blockIdx.y == 0);
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        const double sv = __ldcg(totals + ch), qv = __ldcg(totals + C + ch);
        const double m = sv / count;
        double var = qv / count - m * m;
        if (var < 0.0) var = 0.0;
        const float istd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
        const float sc = g * istd, sh = b - (float)m * g * istd;
        s_tab[ch] = sc;
        s_tab[C + ch] = sh;
        if (first) {
            mean[ch] = (float)m;
            invstd[ch] = istd;
            scale[ch] = sc;
            shift[ch] = sh;
            if (running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
                running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
            }
        }
    }
    __threadfence();                              // the loads of `totals` above are performed before the ticket moves
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1);
    __syncthreads();
    if (s_last) {
        for (int i = threadIdx.x; i < 2 * C; i += 256) totals[i] = 0.0;
        if (threadIdx.x == 0) *ticket = 0;
    }
    const unsigned vpp = (unsigned)C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)W * vpp) return;
    const int w = (int)(idx / vpp);
    const int c = (int)(idx - (unsigned)w * vpp) * 8;
    const float4 s0 = *(const float4*)(s_tab + c), s1 = *(const float4*)(s_tab + c + 4);
    const float4 t0 = *(const float4*)(s_tab + C + c), t1 = *(const float4*)(s_tab + C + c + 4);
    const int Hp = H + 2, Wp = W + 2;
    int cols[3], nc = 0;
    cols[nc++] = w + 1;
    if (w == 1) cols[nc++] = 0;
    if (w == W - 2) cols[nc++] = W + 1;
    const int row_end = min((int)(blockIdx.y + 1) * RB, N * H);
    for (int row = blockIdx.y * RB; row < row_end; row++) {
        const int n = row / H, h = row - n * H;
        const int64_t pix = (int64_t)row * W + w;
        float rr[8];
        load_raw8(raw, pix * C + c, raw_dtype, rr);
        float v[8] = {rr[0] * s0.x + t0.x, rr[1] * s0.y + t0.y, rr[2] * s0.z + t0.z, rr[3] * s0.w + t0.w,
                      rr[4] * s1.x + t1.x, rr[5] * s1.y + t1.y, rr[6] * s1.z + t1.z, rr[7] * s1.w + t1.w};
        float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (drop) {
            const float4 d0 = *(const float4*)(drop + n * C + c), d1 = *(const float4*)(drop + n * C + c + 4);
            dm[0] = d0.x; dm[1] = d0.y; dm[2] = d0.z; dm[3] = d0.w; dm[4] = d1.x; dm[5] = d1.y; dm[6] = d1.z; dm[7] = d1.w;
        }
        __align__(16) __half o[8];
        __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            float z = v[e];
            z = z > 0.f ? z : z * slope;
            z *= dm[e];
            o[e] = __float2half_rn(z);
            ob[e] = __float2bfloat16_rn(z);
        }
        const uint4 ov = *(const uint4*)o;
        const uint4 ovb = *(const uint4*)ob;
        int rows[3], nr = 0;
        rows[nr++] = h + 1;
        if (h == 1) rows[nr++] = 0;
        if (h == H - 2) rows[nr++] = H + 1;
        for (int a = 0; a < nr; a++)
            for (int b = 0; b < nc; b++) {
                const int64_t o_ = (((int64_t)n * Hp + rows[a]) * Wp + cols[b]) * C + c;
                *(uint4*)(act + o_) = ov;
                if (act_b) *(uint4*)(act_b + o_) = ovb;
            }
    }
}

// -------------------------------------------------------------------------------------------
// backward pass 1
// -------------------------------------------------------------------------------------------
struct GSrcs {
    rnr_gsrc_t s[2];
    int n;
};

__device__ __forceinline__ void load8(const void* base, int64_t idx, int dtype, float* out) {
    if (dtype == RNR_F32) {
        const float4 a = *(const float4*)((const float*)base + idx);
        const float4 b = *(const float4*)((const float*)base + idx + 4);
        out[0] += a.x; out[1] += a.y; out[2] += a.z; out[3] += a.w;
        out[4] += b.x; out[5] += b.y; out[6] += b.z; out[7] += b.w;
    } else {
        const uint4 u = *(const uint4*)((const unsigned short*)base + idx);
        const unsigned short* us = (const unsigned short*)&u;
#pragma unroll
        for (int e = 0; e < 8; e++) out[e] += cvt16(us[e], dtype);
    }
}

__global__ void __launch_bounds__(256, 4) bn_bwd_reduce_kernel(const GSrcs srcs, const float* __restrict__ raw,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ drop, float slope,
                                     __nv_bfloat16* __restrict__ gz, float* __restrict__ partials,
                                     int N, int H, int W, int C, int ppb) {
    // block b walks row segments (row = n*H + h, segment = ppb consecutive pixels); thread = (pixel lane pl, 8-channel vector cv).
    // Per-channel constants live in shared memory (not registers) so that 4 CTAs fit per SM; the second moment is accumulated
    // as sum(g * (x - mean)) and scaled by invstd once per block.  Block-level reduction: every thread parks its 16 sums in
    // shared memory [pl][2C] and 2C threads add the ppb rows (conflict-free), no atomics.
    extern __shared__ float smem[];    // [3][C] scale / shift / mean, then [ppb][2C] per-thread sums
    float* s_sc = smem;
    float* s_sh = s_sc + C;
    float* s_mu = s_sh + C;
    float* s_part = s_mu + C;
    const int vpp = C >> 3;
    const int cv = threadIdx.x % vpp, pl = threadIdx.x / vpp;
    const int c = cv * 8;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_sc[i] = scale[i]; s_sh[i] = shift[i]; s_mu[i] = mean[i]; }
    __syncthreads();
    const int Hp = H + 2, Wp = W + 2;
    float sg[8], sgx[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { sg[e] = 0.f; sgx[e] = 0.f; }
    const int segs_per_row = (W + ppb - 1) / ppb;
    const int nseg = N * H * segs_per_row;
    if (pl < ppb)
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int row = seg / segs_per_row;
        const int w = (seg - row * segs_per_row) * ppb + pl;
        if (w >= W) continue;
        const int n = row / H, h = row - n * H;
        const int64_t pix = (int64_t)row * W + w;
        const float4 r0 = __ldcs((const float4*)(raw + pix * C + c));
        const float4 r1 = __ldcs((const float4*)(raw + pix * C + c + 4));
        float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool border = (h == 1) | (h == H - 2) | (w == 1) | (w == W - 2);
        for (int si = 0; si < srcs.n; si++) {
            const rnr_gsrc_t& s = srcs.s[si];
            if (s.fold) {
                const int64_t center = (((int64_t)n * Hp + h + 1) * Wp + w + 1) * s.ld + s.c0 + c;
                load8(s.ptr, center, s.dtype, g);
                if (border) {
                    // reflect halo: halo row 0 mirrors image row 1, halo row H+1 mirrors row H-2 (same for columns)
                    const int dr = (h == 1) ? -(h + 1) : ((h == H - 2) ? 2 : 0);       // padded row offset of the mirrored copy
                    const int dc = (w == 1) ? -(w + 1) : ((w == W - 2) ? 2 : 0);
                    if (dr) load8(s.ptr, center + (int64_t)dr * Wp * s.ld, s.dtype, g);
                    if (dc) load8(s.ptr, center + (int64_t)dc * s.ld, s.dtype, g);
                    if (dr && dc) load8(s.ptr, center + ((int64_t)dr * Wp + dc) * s.ld, s.dtype, g);
                    // (an image of height/width 3 has h == 1 == H-2: both halo rows mirror the same row)
                    if (h == 1 && h == H - 2) {
                        load8(s.ptr, center + (int64_t)2 * Wp * s.ld, s.dtype, g);
                        if (dc) load8(s.ptr, center + ((int64_t)2 * Wp + dc) * s.ld, s.dtype, g);
                    }
                    if (w == 1 && w == W - 2) {
                        load8(s.ptr, center + (int64_t)2 * s.ld, s.dtype, g);
                        if (dr) load8(s.ptr, center + ((int64_t)dr * Wp + 2) * s.ld, s.dtype, g);
                        if (h == 1 && h == H - 2) load8(s.ptr, center + ((int64_t)2 * Wp + 2) * s.ld, s.dtype, g);
                    }
                }
            } else {
                load8(s.ptr, pix * s.ld + s.c0 + c, s.dtype, g);
            }
        }
        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        __align__(16) __nv_bfloat16 o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float z = r[e] * s_sc[c + e] + s_sh[c + e];
            float gg = g[e] * (z > 0.f ? 1.f : slope);
            if (drop) gg *= drop[n * C + c + e];
            sg[e] += gg;
            sgx[e] += gg * (r[e] - s_mu[c + e]);
            o[e] = __float2bfloat16_rn(gg);
        }
        *(uint4*)(gz + (((int64_t)n * Hp + h + 1) * Wp + w + 1) * C + c) = *(const uint4*)o;
    }
    if (pl < ppb) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
            s_part[pl * 2 * C + c + e] = sg[e];
            s_part[pl * 2 * C + C + c + e] = sgx[e];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < ppb; q++) a += s_part[q * 2 * C + i];
        if (i >= C) a *= invstd[i - C];
        partials[(int64_t)blockIdx.x * 2 * C + i] = a;
    }
}

// Same pass with the finalize step fused in and 4 pixels in flight per thread.
//   * one 512-thread block per SM; every thread owns an 8-channel vector and issues the loads of FOUR pixels (raw fp32 + the
//     packed 16-bit gradient sources) before touching any of them: ~100 KB in flight per SM instead of ~25 KB (the one-pixel
//     version ran at 40 % of the HBM roofline, latency-bound);
//   * block partial sums go to a [2C] fp64 accumulator with atomicAdd(double) (148 adds per address; fp64 accumulation of
//     fp32 partials is order-independent to ~1e-16, far below the fp32 result's rounding);
//   * the LAST block to finish (threadfence + ticket) turns the totals into dgamma / dbeta / the fused apply coefficients,
//     re-zeroes the accumulator and re-arms the ticket -- the separate finalize launch disappears.
constexpr int kRUmax = 4;    // pixel vectors in flight per thread (template parameter RU of bn_bwd_reduce_fin_kernel)

__device__ __forceinline__ void acc_packed(const uint4& u, int dtype, float* out) {
    const unsigned short* us = (const unsigned short*)&u;
#pragma unroll
    for (int e = 0; e < 8; e++) out[e] += cvt16(us[e], dtype);
}

// reflect-halo fold of one gradient source at a border pixel (rare: 4 rows / columns of the image) -- kept out of line so
// that its address arithmetic does not inflate the register budget of the streaming loop
__device__ __noinline__ void fold_border(const rnr_gsrc_t s, int64_t center, int h, int w, int H, int W, int Wp, float* g) {
    const int dr = (h == 1) ? -(h + 1) : ((h == H - 2) ? 2 : 0);       // padded row offset of the mirrored copy
    const int dc = (w == 1) ? -(w + 1) : ((w == W - 2) ? 2 : 0);
    if (dr) load8(s.ptr, center + (int64_t)dr * Wp * s.ld, s.dtype, g);
    if (dc) load8(s.ptr, center + (int64_t)dc * s.ld, s.dtype, g);
    if (dr && dc) load8(s.ptr, center + ((int64_t)dr * Wp + dc) * s.ld, s.dtype, g);
    // (an image of height/width 3 has h == 1 == H-2: both halo rows mirror the same row)
    if (h == 1 && h == H - 2) {
        load8(s.ptr, center + (int64_t)2 * Wp * s.ld, s.dtype, g);
        if (dc) load8(s.ptr, center + ((int64_t)2 * Wp + dc) * s.ld, s.dtype, g);
    }
    if (w == 1 && w == W - 2) {
        load8(s.ptr, center + (int64_t)2 * s.ld, s.dtype, g);
        if (dr) load8(s.ptr, center + ((int64_t)dr * Wp + 2) * s.ld, s.dtype, g);
        if (h == 1 && h == H - 2) load8(s.ptr, center + ((int64_t)2 * Wp + 2) * s.ld, s.dtype, g);
    }
}

// pixel index -> (n, h, w); shifts when H*W and W are powers of two (every U-Net level of a 2^k image), divisions otherwise
__device__ __forceinline__ void split_pix(int pix, int HW, int W, int lhw, int lw, int& n, int& h, int& w) {
    if (lhw >= 0) {
        n = pix >> lhw;
        const int rem = pix & (HW - 1);
        h = rem >> lw;
        w = rem & (W - 1);
    } else {
        n = pix / HW;
        const int rem = pix - n * HW;
        h = rem / W;
        w = rem - h * W;
    }
}

template <int NSRC, int kRU>
__global__ void __launch_bounds__(512, 1) bn_bwd_reduce_fin_kernel(const GSrcs srcs, const void* __restrict__ raw, int raw_dtype,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ drop, float slope,
                                     __nv_bfloat16* __restrict__ gz, double* __restrict__ totals, int* __restrict__ ticket,
                                     double count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     const float* __restrict__ gamma, float* __restrict__ coef,
                                     int N, int H, int W, int C, int ppb, int prow, int lhw, int lw) {
    extern __shared__ float smem[];    // [prow][2C] partial rows
    __shared__ int s_last;
    pdl_launch_dependents();
    pdl_wait();
    float* s_part = smem;
    const int vpp = C >> 3;
    const int cv = threadIdx.x % vpp, pl = threadIdx.x / vpp;
    const int c = cv * 8;
    // per-channel constants of this thread's 8 channels stay in registers for the whole kernel (the profile of the version that
    // re-read them from shared memory per pixel was issue-bound: 350 instructions per pixel vector at 16 warps / SM)
    float sc[8], sh[8], mu[8], dr[8];
    {
        const float4 a0 = *(const float4*)(scale + c), a1 = *(const float4*)(scale + c + 4);
        const float4 b0 = *(const float4*)(shift + c), b1 = *(const float4*)(shift + c + 4);
        const float4 m0 = *(const float4*)(mean + c), m1 = *(const float4*)(mean + c + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
        mu[0] = m0.x; mu[1] = m0.y; mu[2] = m0.z; mu[3] = m0.w; mu[4] = m1.x; mu[5] = m1.y; mu[6] = m1.z; mu[7] = m1.w;
#pragma unroll
        for (int e = 0; e < 8; e++) dr[e] = (drop && N == 1) ? drop[c + e] : 1.f;
    }
    const bool drop_per_pixel = drop && N > 1;
    const int Hp = H + 2, Wp = W + 2;
    const int HW = H * W;
    const int P = N * HW;
    float sg[8], sgx[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { sg[e] = 0.f; sgx[e] = 0.f; }
    const int units = (P + ppb - 1) / ppb;
    const rnr_gsrc_t S0 = srcs.s[0];
    const rnr_gsrc_t S1 = srcs.s[NSRC - 1];
    const bool pre0 = S0.dtype != RNR_F32;                              // 16-bit sources are prefetched as packed vectors
    const bool pre1 = NSRC > 1 && S1.dtype != RNR_F32;
    if (pl < ppb)
    for (int u0 = blockIdx.x; u0 < units; u0 += kRU * gridDim.x) {
        float4 r0[kRU], r1[kRU];                     // fp32 raw: 8 floats; 16-bit raw: r0 carries the packed vector
        uint4 q0[kRU], q1[NSRC > 1 ? kRU : 1];
        int pixv[kRU];
        // ---- load phase: everything a pixel needs from HBM, for kRU pixels ----
#pragma unroll
        for (int j = 0; j < kRU; j++) {
            const int u = u0 + j * gridDim.x;
            const int pix = u * ppb + pl;
            pixv[j] = (u < units && pix < P) ? pix : -1;
            if (pixv[j] >= 0) {
                if (raw_dtype == RNR_F32) {
                    r0[j] = __ldcs((const float4*)((const float*)raw + (int64_t)pix * C + c));
                    r1[j] = __ldcs((const float4*)((const float*)raw + (int64_t)pix * C + c + 4));
                } else {
                    r0[j] = __ldcs((const float4*)((const unsigned short*)raw + (int64_t)pix * C + c));
                }
                int n, h, w;
                split_pix(pix, HW, W, lhw, lw, n, h, w);
                const int64_t ctr = ((int64_t)n * Hp + h + 1) * Wp + w + 1;
                if (pre0) q0[j] = *(const uint4*)((const unsigned short*)S0.ptr + (S0.fold ? ctr : (int64_t)pix) * S0.ld + S0.c0 + c);
                if (NSRC > 1 && pre1) q1[j] = *(const uint4*)((const unsigned short*)S1.ptr + (S1.fold ? ctr : (int64_t)pix) * S1.ld + S1.c0 + c);
            }
        }
        // ---- compute phase ----
#pragma unroll
        for (int j = 0; j < kRU; j++) {
            const int pix = pixv[j];
            if (pix < 0) continue;
            int n, h, w;
            split_pix(pix, HW, W, lhw, lw, n, h, w);
            const int64_t ctr = ((int64_t)n * Hp + h + 1) * Wp + w + 1;
            float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const bool border = (h == 1) | (h == H - 2) | (w == 1) | (w == W - 2);
            {
                const int64_t o = (S0.fold ? ctr : (int64_t)pix) * S0.ld + S0.c0 + c;
                if (pre0) acc_packed(q0[j], S0.dtype, g); else load8(S0.ptr, o, S0.dtype, g);
                if (S0.fold && border) fold_border(S0, o, h, w, H, W, Wp, g);
            }
            if (NSRC > 1) {
                const int64_t o = (S1.fold ? ctr : (int64_t)pix) * S1.ld + S1.c0 + c;
                if (pre1) acc_packed(q1[j], S1.dtype, g); else load8(S1.ptr, o, S1.dtype, g);
                if (S1.fold && border) fold_border(S1, o, h, w, H, W, Wp, g);
            }
            float r[8];
            if (raw_dtype == RNR_F32) {
                r[0] = r0[j].x; r[1] = r0[j].y; r[2] = r0[j].z; r[3] = r0[j].w; r[4] = r1[j].x; r[5] = r1[j].y; r[6] = r1[j].z; r[7] = r1[j].w;
            } else {
                const unsigned short* us = (const unsigned short*)&r0[j];
#pragma unroll
                for (int e = 0; e < 8; e++) r[e] = cvt16(us[e], raw_dtype);
            }
            __align__(16) __nv_bfloat16 o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float z = r[e] * sc[e] + sh[e];
                float gg = g[e] * (z > 0.f ? 1.f : slope);
                gg *= drop_per_pixel ? drop[n * C + c + e] : dr[e];
                sg[e] += gg;
                sgx[e] += gg * (r[e] - mu[e]);
                o[e] = __float2bfloat16_rn(gg);
            }
            *(uint4*)(gz + ctr * C + c) = *(const uint4*)o;
        }
    }
    // pixel lanes that share a warp (vpp < 32, a power of two there): butterfly over lane bits >= log2(vpp)
    if (vpp < 32 && (vpp & (vpp - 1)) == 0) {
        for (int off = vpp; off < 32; off <<= 1) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                sg[e] += __shfl_xor_sync(0xffffffffu, sg[e], off);
                sgx[e] += __shfl_xor_sync(0xffffffffu, sgx[e], off);
            }
        }
        const int lane = threadIdx.x & 31;
        if (lane < vpp) {
            const int rowi = threadIdx.x >> 5;
#pragma unroll
            for (int e = 0; e < 8; e++) { s_part[rowi * 2 * C + c + e] = sg[e]; s_part[rowi * 2 * C + C + c + e] = sgx[e]; }
        }
    } else if (pl < ppb) {
#pragma unroll
        for (int e = 0; e < 8; e++) { s_part[pl * 2 * C + c + e] = sg[e]; s_part[pl * 2 * C + C + c + e] = sgx[e]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < prow; q++) a += s_part[q * 2 * C + i];
        if (i >= C) a *= invstd[i - C];
        if (a != 0.f) atomicAdd(totals + i, (double)a);
    }
    // ---- ticket: the last block to arrive finalizes ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        const double sv = __ldcg(totals + ch), qv = __ldcg(totals + C + ch);
        totals[ch] = 0.0;                                   // re-arm for the next launch (stream order / graph replay)
        totals[C + ch] = 0.0;
        if (dbeta) dbeta[ch] = (float)sv;
        if (dgamma) dgamma[ch] = (float)qv;
        if (coef) {
            // gz_out = gamma*invstd*(g - c1 - xhat*c2) = A*g + B*raw + D      (same arithmetic as bn_bwd_finalize_kernel)
            const double gi = (double)gamma[ch] * (double)invstd[ch];
            const double k1 = sv / count, k2 = qv / count;
            coef[ch] = (float)gi;
            coef[C + ch] = (float)(-gi * (double)invstd[ch] * k2);
            coef[2 * C + ch] = (float)(gi * ((double)mean[ch] * (double)invstd[ch] * k2 - k1));
        }
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// BatchNorm backward when the per-channel sums are already in `totals` (accumulated by the data-gradient launches of the
// consumer layers, GStats in conv_internal.cuh): ONE streaming pass
//     gz = A * gg + B * raw + D,   gg = (sum of the gradient sources, reflect halo folded) * drop * lrelu'(raw*scale + shift)
// Every block first turns the totals into the per-channel coefficient table (A, B, D, scale, shift) in shared memory -- one
// channel per thread, two fp64 loads each -- then streams PB rounds of 256 / vpp pixels with one 8-channel vector per thread
// and no per-thread constant registers: occupancy, not registers, keeps HBM busy (the variant that held 40 coefficient registers
// per thread ran at 1.5 TB/s).  Block 0 writes dgamma / dbeta; the last block to have READ the totals (ticket) re-zeroes them.
template <int NSRC>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_src_kernel(const GSrcs srcs, const void* __restrict__ raw, int raw_dtype,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gamma, const float* __restrict__ drop, float slope,
                                     __nv_bfloat16* __restrict__ gz, double* __restrict__ totals, int* __restrict__ ticket,
                                     double inv_count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int N, int H, int W, int C, int lhw, int lw, int PB) {
    extern __shared__ __align__(16) float s_coef[];      // [5][C]: A, B, D, scale, shift
    __shared__ int s_last;
    pdl_launch_dependents();
    pdl_wait();
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        const double sv = __ldcg(totals + ch);
        const double is = (double)invstd[ch];
        const double qv = __ldcg(totals + C + ch) * is;
        const double gi = (double)gamma[ch] * is;
        const double k1 = sv * inv_count, k2 = qv * inv_count;
        s_coef[ch] = (float)gi;
        s_coef[C + ch] = (float)(-gi * is * k2);
        s_coef[2 * C + ch] = (float)(gi * ((double)mean[ch] * is * k2 - k1));
        s_coef[3 * C + ch] = scale[ch];
        s_coef[4 * C + ch] = shift[ch];
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[ch] = (float)sv;
            if (dgamma) dgamma[ch] = (float)qv;
        }
    }
    __threadfence();                              // the loads of `totals` above are performed before the ticket moves
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        for (int i = threadIdx.x; i < 2 * C; i += 256) totals[i] = 0.0;
        if (threadIdx.x == 0) *ticket = 0;
    }
    const int vpp = C >> 3;                       // power of two, <= 256
    const int cv = threadIdx.x & (vpp - 1), pl = threadIdx.x / vpp, ppb = 256 / vpp;
    const int c = cv * 8;
    const int Hp = H + 2, Wp = W + 2;
    const int HW = H * W;
    const int P = N * HW;
    const rnr_gsrc_t S0 = srcs.s[0];
    const rnr_gsrc_t S1 = srcs.s[NSRC - 1];
    const int pix0 = blockIdx.x * PB * ppb + pl;
    const bool pre0 = S0.dtype != RNR_F32;
    const bool pre1 = NSRC > 1 && S1.dtype != RNR_F32;
    for (int j0 = 0; j0 < PB; j0 += 2) {
        // ---- load phase: two pixel vectors (raw + packed 16-bit gradient sources) in flight per thread ----
        uint4 rr[2], q0[2], q1[NSRC > 1 ? 2 : 1];
        int pixv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int pix = pix0 + (j0 + j) * ppb;
            pixv[j] = (j0 + j < PB && pix < P) ? pix : -1;
            if (pixv[j] >= 0) {
                rr[j] = __ldcs((const uint4*)((const unsigned short*)raw + (int64_t)pix * C + c));
                int n, h, w;
                split_pix(pix, HW, W, lhw, lw, n, h, w);
                const int64_t ctr = ((int64_t)n * Hp + h + 1) * Wp + w + 1;
                if (pre0) q0[j] = *(const uint4*)((const unsigned short*)S0.ptr + (S0.fold ? ctr : (int64_t)pix) * S0.ld + S0.c0 + c);
                if (NSRC > 1 && pre1) q1[j] = *(const uint4*)((const unsigned short*)S1.ptr + (S1.fold ? ctr : (int64_t)pix) * S1.ld + S1.c0 + c);
            }
        }
        // ---- compute phase ----
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int pix = pixv[j];
            if (pix < 0) continue;
            int n, h, w;
            split_pix(pix, HW, W, lhw, lw, n, h, w);
            const int64_t ctr = ((int64_t)n * Hp + h + 1) * Wp + w + 1;
            float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const bool border = (h == 1) | (h == H - 2) | (w == 1) | (w == W - 2);
            {
                const int64_t o = (S0.fold ? ctr : (int64_t)pix) * S0.ld + S0.c0 + c;
                if (pre0) acc_packed(q0[j], S0.dtype, g); else load8(S0.ptr, o, S0.dtype, g);
                if (S0.fold && border) {              // (rare; a scratch array keeps g itself in registers)
                    float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    fold_border(S0, o, h, w, H, W, Wp, t);
#pragma unroll
                    for (int e = 0; e < 8; e++) g[e] += t[e];
                }
            }
            if (NSRC > 1) {
                const int64_t o = (S1.fold ? ctr : (int64_t)pix) * S1.ld + S1.c0 + c;
                if (pre1) acc_packed(q1[j], S1.dtype, g); else load8(S1.ptr, o, S1.dtype, g);
                if (S1.fold && border) {
                    float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    fold_border(S1, o, h, w, H, W, Wp, t);
#pragma unroll
                    for (int e = 0; e < 8; e++) g[e] += t[e];
                }
            }
            float dr[8];
            if (drop) {
                const float4 d0 = __ldg((const float4*)(drop + n * C + c)), d1 = __ldg((const float4*)(drop + n * C + c + 4));
                dr[0] = d0.x; dr[1] = d0.y; dr[2] = d0.z; dr[3] = d0.w; dr[4] = d1.x; dr[5] = d1.y; dr[6] = d1.z; dr[7] = d1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; e++) dr[e] = 1.f;
            }
            const unsigned short* us = (const unsigned short*)&rr[j];
            __align__(16) __nv_bfloat16 o[8];
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const float4 A = *(const float4*)(s_coef + c + hh * 4), B = *(const float4*)(s_coef + C + c + hh * 4);
                const float4 D = *(const float4*)(s_coef + 2 * C + c + hh * 4);
                const float4 sc = *(const float4*)(s_coef + 3 * C + c + hh * 4), sh = *(const float4*)(s_coef + 4 * C + c + hh * 4);
                const float Av[4] = {A.x, A.y, A.z, A.w}, Bv[4] = {B.x, B.y, B.z, B.w}, Dv[4] = {D.x, D.y, D.z, D.w};
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int e = hh * 4 + k;
                    const float r = cvt16(us[e], raw_dtype);
                    const float z = r * scv[k] + shv[k];
                    const float gg = g[e] * (z > 0.f ? 1.f : slope) * dr[e];
                    o[e] = __float2bfloat16_rn(Av[k] * gg + Bv[k] * r + Dv[k]);
                }
            }
            *(uint4*)(gz + ctr * C + c) = *(const uint4*)o;
        }
    }
}

__global__ void __launch_bounds__(512) bn_bwd_finalize_kernel(const float* __restrict__ partials, int T, int C, double count,
                                       float* dgamma, float* dbeta, float* c1, float* c2, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ invstd, float* coef) {
    __shared__ double s_s[16][33], s_q[16][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int t = w; t < T; t += 16) {
            s += (double)partials[((int64_t)t * 2 + 0) * C + c];
            q += (double)partials[((int64_t)t * 2 + 1) * C + c];
        }
    }
    s_s[w][lane] = s; s_q[w][lane] = q;
    __syncthreads();
    if (w == 0 && c < C) {
        for (int i = 1; i < 16; i++) { s += s_s[i][lane]; q += s_q[i][lane]; }
        if (dbeta) dbeta[c] = (float)s;
        if (dgamma) dgamma[c] = (float)q;
        if (c1) c1[c] = (float)(s / count);
        if (c2) c2[c] = (float)(q / count);
        if (coef) {
            // gz_out = gamma*invstd*(g - c1 - xhat*c2) = A*g + B*raw + D
            const double gi = (double)gamma[c] * (double)invstd[c];
            const double k1 = s / count, k2 = q / count;
            coef[c] = (float)gi;
            coef[C + c] = (float)(-gi * (double)invstd[c] * k2);
            coef[2 * C + c] = (float)(gi * ((double)mean[c] * (double)invstd[c] * k2 - k1));
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(__nv_bfloat16* __restrict__ gz, const void* __restrict__ raw, int raw_dtype,
                                    const float* __restrict__ coef, int N, int H, int W, int C) {
    // gz <- A*gz + B*raw + D with the per-channel coefficients of bn_bwd_finalize (3 vector loads instead of 5x8 scalars)
    pdl_launch_dependents();
    pdl_wait();
    const unsigned vpp = (unsigned)C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)W * vpp) return;
    const int row = blockIdx.y;
    const int n = row / H, h = row - n * H;
    const int w = (int)(idx / vpp);
    const int c = (int)(idx - (unsigned)w * vpp) * 8;
    const int Hp = H + 2, Wp = W + 2;
    const int64_t pix = (int64_t)row * W + w;
    __nv_bfloat16* gp = gz + (((int64_t)n * Hp + h + 1) * Wp + w + 1) * C + c;
    uint4 u = *(const uint4*)gp;
    __nv_bfloat16* gb = (__nv_bfloat16*)&u;
    float r[8];
    load_raw8(raw, pix * C + c, raw_dtype, r);
    const float4 a0 = *(const float4*)(coef + c), a1 = *(const float4*)(coef + c + 4);
    const float4 b0 = *(const float4*)(coef + C + c), b1 = *(const float4*)(coef + C + c + 4);
    const float4 d0 = *(const float4*)(coef + 2 * C + c), d1 = *(const float4*)(coef + 2 * C + c + 4);
    const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float D[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int e = 0; e < 8; e++) gb[e] = __float2bfloat16_rn(A[e] * __bfloat162float(gb[e]) + B[e] * r[e] + D[e]);
    *(uint4*)gp = u;
}

}  // namespace

extern "C" int rnr_bn_finalize(const float* partials, int T, int ld, int C, double count, const float* gamma,
                               const float* beta, float eps, float* mean, float* invstd, float* scale, float* shift,
                               float* running_mean, float* running_var, float momentum, void* stream) {
    RNR_PDL_LAUNCH(bn_finalize_kernel, rnr_cdiv(C, 32), 512, 0, stream, partials, T, ld, C, count, gamma, beta, eps, mean,
                   invstd, scale, shift, running_mean, running_var, momentum);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_act_fwd(const void* raw, int raw_dtype, const float* scale, const float* shift, const float* drop, float slope,
                              void* act, void* act_bf16, int N, int H, int W, int C, void* stream) {
    RNR_REQUIRE(C % 8 == 0, "rnr_bn_act_fwd: C=%d must be a multiple of 8", C);
    RNR_REQUIRE(H >= 2 && W >= 2, "rnr_bn_act_fwd: reflect halo needs H,W >= 2");
    RNR_REQUIRE((int64_t)N * H <= 65535, "rnr_bn_act_fwd: N*H=%lld exceeds the grid limit", (long long)N * H);
    dim3 grid(rnr_cdiv((int64_t)W * (C / 8), 256), N * H);
    RNR_PDL_LAUNCH(bn_act_fwd_kernel, grid, 256, 0, stream, raw, raw_dtype, scale, shift, drop, slope, (__half*)act, (__nv_bfloat16*)act_bf16, N, H, W, C);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_act_fwd_tot(const void* raw, int raw_dtype, double* totals, int* ticket, double count, const float* gamma,
                                  const float* beta, float eps, float* mean, float* invstd, float* scale, float* shift,
                                  float* running_mean, float* running_var, float momentum, const float* drop, float slope, void* act,
                                  void* act_bf16, int N, int H, int W, int C, void* stream) {
    RNR_REQUIRE(C % 8 == 0 && C <= 4096, "rnr_bn_act_fwd_tot: C=%d must be a multiple of 8 (<= 4096)", C);
    RNR_REQUIRE(H >= 2 && W >= 2, "rnr_bn_act_fwd_tot: reflect halo needs H,W >= 2");
    RNR_REQUIRE(totals && ticket && mean && invstd && scale && shift, "rnr_bn_act_fwd_tot: null pointer");
    const int bx = rnr_cdiv((int64_t)W * (C / 8), 256);
    // rows per block: ~2000 blocks for the big layers (one ticket atomic + one coefficient prologue per block), >= 1 row
    int RB = 1;
    while (RB < 8 && (int64_t)bx * rnr_cdiv((int64_t)N * H, RB) > 2048) RB *= 2;
    dim3 grid(bx, rnr_cdiv((int64_t)N * H, RB));
    RNR_REQUIRE(grid.y <= 65535, "rnr_bn_act_fwd_tot: N*H=%lld exceeds the grid limit", (long long)N * H);
    RNR_PDL_LAUNCH(bn_act_fwd_tot_kernel, grid, 256, (size_t)2 * C * sizeof(float), stream, raw, raw_dtype, totals, ticket, count, gamma, beta,
                   eps, mean, invstd, scale, shift, running_mean, running_var, momentum, drop, slope, (__half*)act,
                   (__nv_bfloat16*)act_bf16, N, H, W, C, RB);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_bwd_reduce(const rnr_gsrc_t* srcs, int nsrc, const float* raw, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, const float* drop, float slope, void* gz,
                                 float* partials, int* T_out, int N, int H, int W, int C, void* stream) {
    RNR_REQUIRE(C % 8 == 0 && C <= 2048, "rnr_bn_bwd_reduce: bad C=%d", C);
    RNR_REQUIRE(nsrc >= 1 && nsrc <= 2, "rnr_bn_bwd_reduce: nsrc must be 1 or 2");
    GSrcs gs;
    gs.n = nsrc;
    for (int i = 0; i < nsrc; i++) gs.s[i] = srcs[i];
    const int vpp = C / 8;
    int ppb = 256 / vpp;
    if (ppb < 1) ppb = 1;
    const int threads = vpp * ppb > 256 ? 256 : vpp * ppb;   // vpp<=256 guaranteed by C<=2048
    const int64_t nseg = (int64_t)N * H * rnr_cdiv(W, ppb);
    int T = rnr_cdiv(nseg, 4);
    if (T > 148 * 4) T = 148 * 4;
    if (T < 1) T = 1;
    if (T_out) *T_out = T;
    bn_bwd_reduce_kernel<<<T, threads, (3 * C + (size_t)ppb * 2 * C) * sizeof(float), (cudaStream_t)stream>>>(
        gs, raw, scale, shift, mean, invstd, drop, slope, (__nv_bfloat16*)gz, partials, N, H, W, C, ppb);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_bwd_reduce_fin(const rnr_gsrc_t* srcs, int nsrc, const void* raw, int raw_dtype, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, const float* drop, float slope, void* gz,
                                     double* totals, int* ticket, double count, float* dgamma, float* dbeta, const float* gamma,
                                     float* coef, int N, int H, int W, int C, void* stream) {
    RNR_REQUIRE(C % 8 == 0 && C <= 2048, "rnr_bn_bwd_reduce_fin: bad C=%d", C);
    RNR_REQUIRE(nsrc >= 1 && nsrc <= 2, "rnr_bn_bwd_reduce_fin: nsrc must be 1 or 2");
    RNR_REQUIRE(ticket && totals, "rnr_bn_bwd_reduce_fin: ticket / totals required");
    RNR_REQUIRE(!coef || gamma, "rnr_bn_bwd_reduce_fin: coef needs gamma");
    RNR_REQUIRE((int64_t)N * H * W < (1ll << 31), "rnr_bn_bwd_reduce_fin: too many pixels");
    GSrcs gs;
    gs.n = nsrc;
    for (int i = 0; i < nsrc; i++) gs.s[i] = srcs[i];
    const int vpp = C / 8;
    const int64_t P = (int64_t)N * H * W;
    int ppb = 512 / vpp;
    if (ppb < 1) ppb = 1;
    const bool pow2 = (vpp & (vpp - 1)) == 0;
    while (ppb > 1 && (int64_t)(ppb / 2) >= P && pow2) ppb /= 2;        // tiny layers: do not launch idle pixel lanes
    if (!pow2 && ppb > P) ppb = (int)P;
    int threads = vpp * ppb;
    threads = (threads + 31) / 32 * 32;
    RNR_REQUIRE(threads <= 512, "rnr_bn_bwd_reduce_fin: C=%d needs %d threads", C, threads);
    const bool shuffle = vpp < 32 && pow2;
    const int prow = shuffle ? threads / 32 : ppb;
    const int64_t units = (P + ppb - 1) / ppb;
    int T = (int)(units < 148 ? units : 148);
    if (T < 1) T = 1;
    const size_t smem = (size_t)prow * 2 * C * sizeof(float);
    int lhw = -1, lw = -1;
    {
        const int HW = H * W;
        if ((HW & (HW - 1)) == 0 && (W & (W - 1)) == 0) {
            lhw = 0; while ((1 << lhw) < HW) lhw++;
            lw = 0; while ((1 << lw) < W) lw++;
        }
    }
    RNR_REQUIRE(smem <= 160 * 1024, "rnr_bn_bwd_reduce_fin: C=%d needs %zu bytes of shared memory", C, smem);
    // pixel vectors in flight per thread.  Measured on B200 (bench.py, 512^2 step): RU = 2 / 3 / 4 -> 201.7 / 200.2 / 198.5 views/s --
    // the extra loads in flight of RU > 2 do not pay for the register spills of a 128-register, 512-thread block.  RNR_BN_RU overrides.
    int ru = 2;
    { const char* e = getenv("RNR_BN_RU"); if (e && atoi(e) >= 2 && atoi(e) <= kRUmax) ru = atoi(e); }
#define RNR_BN_LAUNCH(NS, RU)                                                                                                        \
    do {                                                                                                                             \
        RNR_ONCE_PER_DEVICE({ RNR_CHECK(cudaFuncSetAttribute(bn_bwd_reduce_fin_kernel<NS, RU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); }); \
        RNR_PDL_LAUNCH((bn_bwd_reduce_fin_kernel<NS, RU>), T, threads, smem, stream,                                                \
            gs, raw, raw_dtype, scale, shift, mean, invstd, drop, slope, (__nv_bfloat16*)gz, totals, ticket, count, dgamma, dbeta,   \
            gamma, coef, N, H, W, C, ppb, prow, lhw, lw);                                                                            \
    } while (0)
    if (nsrc == 1) { if (ru == 2) RNR_BN_LAUNCH(1, 2); else if (ru == 3) RNR_BN_LAUNCH(1, 3); else RNR_BN_LAUNCH(1, 4); }
    else { if (ru == 2) RNR_BN_LAUNCH(2, 2); else if (ru == 3) RNR_BN_LAUNCH(2, 3); else RNR_BN_LAUNCH(2, 4); }
#undef RNR_BN_LAUNCH
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_bwd_apply_src(const rnr_gsrc_t* srcs, int nsrc, const void* raw, int raw_dtype, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* gamma, const float* drop,
                                    float slope, void* gz, double* totals, int* ticket, double count, float* dgamma, float* dbeta,
                                    int N, int H, int W, int C, void* stream) {
    const int vpp = C / 8;
    RNR_REQUIRE(C % 8 == 0 && vpp >= 1 && vpp <= 256 && (vpp & (vpp - 1)) == 0, "rnr_bn_bwd_apply_src: C=%d must be 8 x a power of two <= 2048", C);
    RNR_REQUIRE(nsrc >= 1 && nsrc <= 2, "rnr_bn_bwd_apply_src: nsrc must be 1 or 2");
    RNR_REQUIRE(raw_dtype != RNR_F32, "rnr_bn_bwd_apply_src: 16-bit raw tensor required");
    RNR_REQUIRE(ticket && totals && gamma, "rnr_bn_bwd_apply_src: ticket / totals / gamma required");
    RNR_REQUIRE((int64_t)N * H * W < (1ll << 31), "rnr_bn_bwd_apply_src: too many pixels");
    GSrcs gs;
    gs.n = nsrc;
    for (int i = 0; i < nsrc; i++) gs.s[i] = srcs[i];
    const int ppb = 256 / vpp;
    const int64_t P = (int64_t)N * H * W;
    const int64_t rounds = (P + ppb - 1) / ppb;             // pixel rounds of one block-wide sweep
    // PB rounds per block: enough blocks to fill the machine several times over, few enough that the per-block coefficient
    // prologue (2 fp64 loads per channel, one ticket atomic) stays a small fraction of the block's work
    int PB = 8;
    while (PB > 2 && rounds / PB < 148 * 8) PB /= 2;
    const int blocks = (int)((rounds + PB - 1) / PB);
    const size_t smem = (size_t)5 * C * sizeof(float);
    int lhw = -1, lw = -1;
    {
        const int HW = H * W;
        if ((HW & (HW - 1)) == 0 && (W & (W - 1)) == 0) {
            lhw = 0; while ((1 << lhw) < HW) lhw++;
            lw = 0; while ((1 << lw) < W) lw++;
        }
    }
    if (nsrc == 1)
        RNR_PDL_LAUNCH(bn_bwd_apply_src_kernel<1>, blocks, 256, smem, stream, gs, raw, raw_dtype, scale, shift, mean, invstd, gamma, drop, slope,
                                                                             (__nv_bfloat16*)gz, totals, ticket, 1.0 / count, dgamma, dbeta,
                                                                             N, H, W, C, lhw, lw, PB);
    else
        RNR_PDL_LAUNCH(bn_bwd_apply_src_kernel<2>, blocks, 256, smem, stream, gs, raw, raw_dtype, scale, shift, mean, invstd, gamma, drop, slope,
                                                                             (__nv_bfloat16*)gz, totals, ticket, 1.0 / count, dgamma, dbeta,
                                                                             N, H, W, C, lhw, lw, PB);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_bwd_finalize(const float* partials, int T, int C, double count, float* dgamma, float* dbeta, float* c1,
                                   float* c2, const float* gamma, const float* mean, const float* invstd, float* coef,
                                   void* stream) {
    RNR_REQUIRE(!coef || (gamma && mean && invstd), "rnr_bn_bwd_finalize: coef needs gamma / mean / invstd");
    bn_bwd_finalize_kernel<<<rnr_cdiv(C, 32), 512, 0, (cudaStream_t)stream>>>(partials, T, C, count, dgamma, dbeta, c1, c2, gamma, mean,
                                                                            invstd, coef);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bn_bwd_apply(void* gz, const void* raw, int raw_dtype, const float* coef, int N, int H, int W, int C, void* stream) {
    RNR_REQUIRE(C % 8 == 0, "rnr_bn_bwd_apply: C=%d must be a multiple of 8", C);
    RNR_REQUIRE((int64_t)N * H <= 65535, "rnr_bn_bwd_apply: N*H=%lld exceeds the grid limit", (long long)N * H);
    dim3 grid(rnr_cdiv((int64_t)W * (C / 8), 256), N * H);
    RNR_PDL_LAUNCH(bn_bwd_apply_kernel, grid, 256, 0, stream, (__nv_bfloat16*)gz, raw, raw_dtype, coef, N, H, W, C);
    RNR_LAUNCH_CHECK();
    return 0;
}
