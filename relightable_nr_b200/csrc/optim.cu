// Optimiser kernels of the fused training step (train_rnr.py:376,618-623: torch.optim.Adam(lr) over the U-Net, the neural
// textures and the SH coefficients).  The reference's optimiser is three passes over ~45 M weights in this design: un-transpose
// the weight gradients from GEMM order, Adam, re-derive the 16-bit GEMM matrices.  Here:
//
//   adam_wunpack_kernel   gradient scratch [tap][co][ci] (GEMM order, what the tcgen05 weight-gradient kernels accumulate)
//                         -> Adam on the fp32 master weight + m + v in parameter layout, in ONE pass: the un-transposed
//                         gradient never exists in HBM, and the scratch is re-zeroed on the way (no memset next step).
//                         32 B/parameter instead of 8 (un-transpose) + 28 (Adam) + 4 (memset).
//   adam_multi_kernel     one launch for all the small tensors (biases, BatchNorm affine, 4 texture levels, SH coefficients)
//                         through a job table; optionally re-zeroes the gradients; the LAST block advances the step counter.
//   loss_combine_kernel / dropout_mask_kernel   the scalar glue and the Dropout2d channel masks without ATen launches.
//
// Arithmetic follows torch's fused Adam exactly (exp_avg = lerp(exp_avg, g, 1-beta1); exp_avg_sq = beta2 v + (1-beta2) g^2;
// denom = sqrt(v)/sqrt(1-beta2^t) + eps; p -= lr/(1-beta1^t) * m/denom), bias corrections from a device-resident step
// counter so that the whole optimiser lives inside a CUDA graph.
#include "common.cuh"
#include <vector>

namespace {

constexpr int CB = 64, MAXT = 16, UCO = 8;

struct AdamHyper {
    float lr, beta1, beta2, eps, gscale;
};

__device__ __forceinline__ void bias_corrections(const float* __restrict__ step, const AdamHyper& h, float& step_size, float& bc2_sqrt) {
    const double t = (double)(*step) + 1.0;
    const double bc1 = 1.0 - pow((double)h.beta1, t);
    const double bc2 = 1.0 - pow((double)h.beta2, t);
    step_size = (float)((double)h.lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h, float step_size, float bc2_sqrt) {
    m = m + (1.f - h.beta1) * (g - m);
    v = h.beta2 * v + (1.f - h.beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + h.eps;
    p -= step_size * (m / denom);
}

struct WMat {
    void* base;
    int64_t ld;
    int32_t dtype, ntaps, sub_rows, r0, r1;
    int8_t inv[16];
};

struct AdamWJob {
    float* scratch;        // [ntaps][cout][cin]
    float* p;              // parameter layout: element (co, ci, t) at co*s_co + ci*s_ci + t
    float* m;
    float* v;
    float* gdst;           // optional: the un-transposed gradient (parameter layout); nullptr = not materialised
    int32_t cout, cin, ntaps, tiles_c;
    int64_t s_co, s_ci;
    int32_t blk0, nblk;
    WMat fwd, dgrad;       // 16-bit GEMM matrices of the NEXT step, written from the updated weight (base nullptr: not here)
};

// Block = 8 output channels x one 64-input-channel chunk x all taps of one conv weight.
//   1. gradient tile from the GEMM-order scratch (re-zeroed on the way)            -> shared memory [co][tap][ci]
//   2. Adam in parameter layout (contiguous runs of p / m / v); the updated weight replaces the gradient in the tile
//   3. forward GEMM matrix:       row co, columns [ci chunk][tap slot][64 ci]      -> 128-byte runs, 16-byte stores
//   4. data-gradient GEMM matrix: row ci, columns [co chunk][tap slot][64 co]      -> the block's 8 co = one 16-byte store per (ci, tap)
__global__ void __launch_bounds__(256) adam_wunpack_kernel(const AdamWJob* __restrict__ jobs, const int* __restrict__ blk2job,
                                                         const float* __restrict__ step, const AdamHyper h, int zero_src) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[UCO][MAXT][CB + 1];
    const AdamWJob& J = jobs[blk2job[blockIdx.x]];
    const int local = blockIdx.x - J.blk0;
    const int cog = local / J.tiles_c, c0 = (local - cog * J.tiles_c) * CB;
    const int rq = threadIdx.x >> 6, cc = threadIdx.x & 63;    // thread = (co row of the block modulo 4, input channel of the chunk)
    const int nc = min(CB, J.cin - c0);
    const int ntaps = J.ntaps;
    float step_size, bc2_sqrt;
    bias_corrections(step, h, step_size, bc2_sqrt);
    // ---- 1. gradient tile ----
#pragma unroll
    for (int half = 0; half < UCO / 4; half++) {
        const int r = rq + 4 * half, co = cog * UCO + r;
        if (co < J.cout && cc < nc) {
            float* sp = J.scratch + (int64_t)co * J.cin + c0 + cc;
            const int64_t tstride = (int64_t)J.cout * J.cin;
            float g[MAXT];
#pragma unroll
            for (int t = 0; t < MAXT; t++) g[t] = (t < ntaps) ? __ldcs(sp + t * tstride) : 0.f;
#pragma unroll
            for (int t = 0; t < MAXT; t++)
                if (t < ntaps) {
                    tile[r][t][cc] = g[t] * h.gscale;
                    if (zero_src) __stcs(sp + t * tstride, 0.f);
                }
        } else {
#pragma unroll
            for (int t = 0; t < MAXT; t++)
                if (t < ntaps) tile[r][t][cc] = 0.f;          // padding rows / channels read as zero weights below
        }
    }
    __syncthreads();
    // ---- 2. Adam in parameter layout ----
    const unsigned magic = (65536u + ntaps - 1) / ntaps;       // j / ntaps for j < 1024, ntaps <= 16: exact
#pragma unroll
    for (int half = 0; half < UCO / 4; half++) {
        const int r = rq + 4 * half, co = cog * UCO + r;
        if (co < J.cout) {
            const int64_t base = (int64_t)co * J.s_co + (int64_t)c0 * J.s_ci;
            const int run = nc * ntaps;
            for (int j = cc; j < run; j += 64) {
                const int ci = (int)(((unsigned)j * magic) >> 16), t = j - ci * ntaps;
                const int64_t idx = base + (int64_t)ci * J.s_ci + t;
                const float g = tile[r][t][ci];
                float p = J.p[idx], m = J.m[idx], v = J.v[idx];
                adam_update(p, g, m, v, h, step_size, bc2_sqrt);
                J.p[idx] = p; J.m[idx] = m; J.v[idx] = v;
                if (J.gdst) J.gdst[idx] = g;
                tile[r][t][ci] = p;                            // (this thread is the only reader / writer of the element)
            }
        }
    }
    if (!J.fwd.base && !J.dgrad.base) return;
    __syncthreads();
    // ---- 3. forward matrix ----
    if (J.fwd.base) {
        unsigned short* W = (unsigned short*)J.fwd.base;
        const int items = UCO * ntaps * 8;                     // (co row, source tap, group of 8 input channels)
        for (int i = threadIdx.x; i < items; i += 256) {
            const int cg = i & 7, t = (i >> 3) % ntaps, r = (i >> 3) / ntaps;
            const int co = cog * UCO + r, slot = J.fwd.inv[t];
            if (co >= J.cout || slot < 0) continue;
            __align__(16) unsigned short o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = f2b16(tile[r][t][cg * 8 + e], J.fwd.dtype);   // (channels >= cin hold zeros)
            const int sidx = slot >> 4, k = slot & 15;
            *(uint4*)(W + ((int64_t)sidx * J.fwd.sub_rows + co) * J.fwd.ld + ((int64_t)(c0 >> 6) * J.fwd.ntaps + k) * 64 + cg * 8) = *(const uint4*)o;
        }
    }
    // ---- 4. data-gradient matrix ----
    if (J.dgrad.base) {
        unsigned short* W = (unsigned short*)J.dgrad.base;
        const int co0 = cog * UCO;
        const int items = CB * ntaps;                          // (input channel, source tap)
        for (int i = threadIdx.x; i < items; i += 256) {
            const int ci = i & 63, t = i >> 6;
            const int gci = c0 + ci, slot = J.dgrad.inv[t];
            if (gci >= J.cin || gci < J.dgrad.r0 || gci >= J.dgrad.r1 || slot < 0) continue;
            __align__(16) unsigned short o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = f2b16(tile[e][t][ci], J.dgrad.dtype);         // (rows >= cout hold zeros)
            const int sidx = slot >> 4, k = slot & 15;
            *(uint4*)(W + ((int64_t)sidx * J.dgrad.sub_rows + (gci - J.dgrad.r0)) * J.dgrad.ld + ((int64_t)(co0 >> 6) * J.dgrad.ntaps + k) * 64 +
                      (co0 & 63)) = *(const uint4*)o;
        }
    }
}

struct AdamJob {
    float* p;
    float* g;
    float* m;
    float* v;
    int64_t n;
    int32_t blk0, nblk;
};

constexpr int kAdamChunk = 2048;     // elements per block: 256 threads x 2 x float4

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamJob* __restrict__ jobs, const int* __restrict__ blk2job, float* step,
                                                       int* ticket, const AdamHyper h, int zero_grad, int advance_step) {
    pdl_launch_dependents();
    pdl_wait();
    const AdamJob& J = jobs[blk2job[blockIdx.x]];
    const int64_t e0 = (int64_t)(blockIdx.x - J.blk0) * kAdamChunk;
    const int64_t n = min((int64_t)kAdamChunk, J.n - e0);
    float step_size, bc2_sqrt;
    bias_corrections(step, h, step_size, bc2_sqrt);
    float* p = J.p + e0; float* g = J.g + e0; float* m = J.m + e0; float* v = J.v + e0;
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = threadIdx.x; i < n4; i += 256) {
            float4 pp = ((float4*)p)[i], gg = ((const float4*)g)[i], mm = ((float4*)m)[i], vv = ((float4*)v)[i];
            float* P = (float*)&pp; float* G = (float*)&gg; float* M = (float*)&mm; float* V = (float*)&vv;
#pragma unroll
            for (int e = 0; e < 4; e++) adam_update(P[e], G[e] * h.gscale, M[e], V[e], h, step_size, bc2_sqrt);
            ((float4*)p)[i] = pp; ((float4*)m)[i] = mm; ((float4*)v)[i] = vv;
            if (zero_grad) ((float4*)g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
            float pp = p[i], mm = m[i], vv = v[i];
            adam_update(pp, g[i] * h.gscale, mm, vv, h, step_size, bc2_sqrt);
            p[i] = pp; m[i] = mm; v[i] = vv;
            if (zero_grad) g[i] = 0.f;
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += 256) {
            float pp = p[i], mm = m[i], vv = v[i];
            adam_update(pp, g[i] * h.gscale, mm, vv, h, step_size, bc2_sqrt);
            p[i] = pp; m[i] = mm; v[i] = vv;
            if (zero_grad) g[i] = 0.f;
        }
    }
    if (advance_step) {
        // every block has read *step above; the last one to get here advances it (and re-arms the ticket)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
                *step = *step + 1.f;
                *ticket = 0;
            }
        }
    }
}

// loss = sums[2]/cnt + sums[0]/sums[1]/R * w_chrom + extra[0] + extra[1] ...   (train_rnr.py:608: loss_g = loss_lighting + loss_rn +
// loss_rays_lt_chrom + loss_alb), from the device-side accumulators of the tail / small-loss kernels
__global__ void loss_combine_kernel(double* __restrict__ sums, double cnt, double R, double w_chrom, const double* __restrict__ extra,
                                    int n_extra, float* __restrict__ out, int n_clear) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double l = sums[2] / cnt + (sums[1] != 0.0 ? sums[0] / sums[1] / R * w_chrom : 0.0);
        for (int i = 0; i < n_extra; i++) l += extra[i];
        out[0] = (float)l;
        for (int i = 0; i < n_clear; i++) sums[i] = 0.0;       // the accumulators are ready for the next step (no memset / fill launch)
    }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// nn.Dropout2d channel masks [n]: 0 with probability p, else 1/(1-p).  Counter-based generator (seed, launch counter, index):
// every launch draws fresh masks without host involvement (CUDA-graph replays included); one block.
__global__ void __launch_bounds__(1024) dropout_mask_kernel(float* __restrict__ out, int n, float p, unsigned long long seed,
                                                          unsigned long long* counter) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned long long s_ctr;
    if (threadIdx.x == 0) s_ctr = *counter;
    __syncthreads();
    const unsigned long long base = splitmix64(seed ^ splitmix64(s_ctr));
    const float keep = 1.f / (1.f - p);
    for (int i = threadIdx.x; i < n; i += 1024) {
        const unsigned long long r = splitmix64(base + (unsigned long long)i);
        const float u = (float)(r >> 40) * (1.f / 16777216.f);       // 24 random bits -> [0, 1)
        out[i] = (u >= p) ? keep : 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = s_ctr + 1ull;
}

template <class Job>
int upload_jobs(const std::vector<Job>& h, int nblocks, Job** d_jobs, int** d_blk2job) {
    RNR_CHECK(cudaMalloc(d_jobs, sizeof(Job) * h.size()));
    RNR_CHECK(cudaMemcpy(*d_jobs, h.data(), sizeof(Job) * h.size(), cudaMemcpyHostToDevice));
    std::vector<int> b2j(nblocks);
    for (size_t i = 0; i < h.size(); i++)
        for (int b = 0; b < h[i].nblk; b++) b2j[h[i].blk0 + b] = (int)i;
    RNR_CHECK(cudaMalloc(d_blk2job, sizeof(int) * nblocks));
    RNR_CHECK(cudaMemcpy(*d_blk2job, b2j.data(), sizeof(int) * nblocks, cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace

struct rnr_adam_plan {
    AdamWJob* d_wjobs = nullptr;
    int* d_wblk = nullptr;
    int wblocks = 0;
    AdamJob* d_jobs = nullptr;
    int* d_blk = nullptr;
    int blocks = 0;
    int* d_ticket = nullptr;
};

extern "C" int rnr_adam_plan_create(const rnr_adam_wjob_t* wjobs, int n_wjobs, const rnr_adam_job_t* jobs, int n_jobs, rnr_adam_plan_t** out) {
    RNR_REQUIRE(out && (n_wjobs + n_jobs) >= 1, "rnr_adam_plan_create: no jobs");
    rnr_adam_plan* p = new rnr_adam_plan();
    if (n_wjobs > 0) {
        std::vector<AdamWJob> h(n_wjobs);
        int blk = 0;
        for (int i = 0; i < n_wjobs; i++) {
            const rnr_adam_wjob_t& s = wjobs[i];
            RNR_REQUIRE(s.ntaps >= 1 && s.ntaps <= MAXT, "adam plan: 1..%d taps, got %d", MAXT, s.ntaps);
            RNR_REQUIRE(s.scratch && s.p && s.m && s.v, "adam plan: null buffer in weight job %d", i);
            AdamWJob& d = h[i];
            d.scratch = s.scratch; d.p = s.p; d.m = s.m; d.v = s.v; d.gdst = s.gdst;
            d.cout = s.cout; d.cin = s.cin; d.ntaps = s.ntaps; d.s_co = s.s_co; d.s_ci = s.s_ci;
            const rnr_wmat_t* src[2] = {&s.fwd, &s.dgrad};
            WMat* dst[2] = {&d.fwd, &d.dgrad};
            for (int f = 0; f < 2; f++) {
                memset(dst[f], 0, sizeof(WMat));
                if (!src[f]->base) continue;
                RNR_REQUIRE(src[f]->dtype == RNR_F16 || src[f]->dtype == RNR_BF16, "adam plan: GEMM matrices must be 16-bit");
                RNR_REQUIRE(src[f]->ld % 64 == 0 && ((uintptr_t)src[f]->base & 15) == 0 && src[f]->ntaps >= 1 && src[f]->ntaps <= MAXT,
                            "adam plan: bad GEMM matrix descriptor (ld %lld, ntaps %d)", (long long)src[f]->ld, src[f]->ntaps);
                dst[f]->base = src[f]->base; dst[f]->ld = src[f]->ld; dst[f]->dtype = src[f]->dtype; dst[f]->ntaps = src[f]->ntaps;
                dst[f]->sub_rows = src[f]->sub_rows; dst[f]->r0 = src[f]->r0; dst[f]->r1 = src[f]->r1;
                for (int t = 0; t < 16; t++) {
                    dst[f]->inv[t] = src[f]->inv[t];
                    RNR_REQUIRE(src[f]->inv[t] < 0 || (src[f]->inv[t] & 15) < src[f]->ntaps, "adam plan: tap slot out of range");
                }
            }
            if (!d.dgrad.base) { d.dgrad.r0 = 0; d.dgrad.r1 = 0; }
            d.tiles_c = rnr_cdiv(s.cin, CB);
            d.blk0 = blk;
            d.nblk = rnr_cdiv(s.cout, UCO) * d.tiles_c;
            blk += d.nblk;
        }
        p->wblocks = blk;
        int rc = upload_jobs(h, blk, &p->d_wjobs, &p->d_wblk);
        if (rc) return rc;
    }
    if (n_jobs > 0) {
        std::vector<AdamJob> h(n_jobs);
        int blk = 0;
        for (int i = 0; i < n_jobs; i++) {
            const rnr_adam_job_t& s = jobs[i];
            RNR_REQUIRE(s.p && s.g && s.m && s.v && s.n >= 1, "adam plan: bad tensor job %d", i);
            AdamJob& d = h[i];
            d.p = s.p; d.g = s.g; d.m = s.m; d.v = s.v; d.n = s.n;
            d.blk0 = blk;
            d.nblk = rnr_cdiv(s.n, kAdamChunk);
            blk += d.nblk;
        }
        p->blocks = blk;
        int rc = upload_jobs(h, blk, &p->d_jobs, &p->d_blk);
        if (rc) return rc;
    }
    RNR_CHECK(cudaMalloc(&p->d_ticket, sizeof(int)));
    RNR_CHECK(cudaMemset(p->d_ticket, 0, sizeof(int)));
    *out = p;
    return 0;
}

extern "C" void rnr_adam_plan_destroy(rnr_adam_plan_t* p) {
    if (!p) return;
    cudaFree(p->d_wjobs); cudaFree(p->d_wblk); cudaFree(p->d_jobs); cudaFree(p->d_blk); cudaFree(p->d_ticket);
    delete p;
}

/* `step`: device float holding the number of optimiser steps taken so far (the kernels use step + 1 for the bias corrections).
 * The conv-weight jobs run first, then the plain tensor jobs; with `advance_step` the last block of the tensor-job launch
 * increments *step (a plan that advances the step must therefore have at least one tensor job). */
extern "C" int rnr_adam_run(const rnr_adam_plan_t* p, float* step, float lr, float beta1, float beta2, float eps, float gscale,
                            int zero_grad, int advance_step, void* stream) {
    RNR_REQUIRE(p && step, "rnr_adam_run: null argument");
    RNR_REQUIRE(!advance_step || p->blocks > 0, "rnr_adam_run: advance_step needs at least one tensor job in the plan");
    AdamHyper h{lr, beta1, beta2, eps, gscale};
    if (p->wblocks > 0) {
        RNR_PDL_LAUNCH(adam_wunpack_kernel, p->wblocks, 256, 0, stream, p->d_wjobs, p->d_wblk, step, h, zero_grad);
        RNR_LAUNCH_CHECK();
    }
    if (p->blocks > 0) {
        RNR_PDL_LAUNCH(adam_multi_kernel, p->blocks, 256, 0, stream, p->d_jobs, p->d_blk, step, p->d_ticket, h, zero_grad, advance_step);
        RNR_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int rnr_loss_combine(double* sums, double cnt, double R, double w_chrom, const double* extra, int n_extra, float* out,
                                int n_clear, void* stream) {
    RNR_REQUIRE(sums && out && cnt > 0 && R > 0, "rnr_loss_combine: bad arguments");
    RNR_PDL_LAUNCH(loss_combine_kernel, 1, 32, 0, stream, sums, cnt, R, w_chrom, extra, n_extra, out, n_clear);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_dropout_masks(float* out, int n, float p, unsigned long long seed, unsigned long long* counter, void* stream) {
    RNR_REQUIRE(out && counter && n >= 1 && p >= 0.f && p < 1.f, "rnr_dropout_masks: bad arguments");
    RNR_PDL_LAUNCH(dropout_mask_kernel, 1, 1024, 0, stream, out, n, p, seed, counter);
    RNR_LAUNCH_CHECK();
    return 0;
}
