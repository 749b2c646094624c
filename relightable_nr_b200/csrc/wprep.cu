// Batched weight preparation: fp32 master weights -> the 16-bit K-major GEMM matrices of every layer, ONE launch.
//
// The U-Net has 22 live layers; each needs a forward matrix (4 for a ConvTranspose: one per output parity) and a
// data-gradient matrix (4 for a stride-2 conv) re-derived from the fp32 parameters after every optimiser step:
// 74 small strided transposes.  A plan holds the job table in device memory; one persistent-style launch walks
// (job, 8-row x 64-column tile) work items.  Each tile is staged through shared memory so that both the fp32 source
// (contiguous along whichever of row/column is the inner dimension of the parameter tensor) and the 16-bit
// destination (contiguous along the column = input-channel dimension) move in full sectors.
//   dst[r*ntaps*cpad + t*cpad + c] = src[r*s_r + c*s_c + tapoff[t]]     (0 for r >= nr or c >= nc)
// or, chunk-major (all taps of a 64-channel chunk adjacent in K, the order the halo kernel consumes):
//   dst[r*ntaps*cpad + ((c/64)*ntaps + t)*64 + c%64]
// Replaces the per-layer weight reshapes that cuDNN does inside nn.Conv2d / nn.ConvTranspose2d
// (pytorch_prototyping/pytorch_prototyping.py:112-115,155-160,242-264).
#include "common.cuh"
#include <vector>

namespace {

constexpr int RB = 8, CB = 64, MAXT = 16;

struct WJob {
    const float* src;
    void* dst;
    int32_t dtype, nr, nr_pad, nc, cpad, ntaps, ts, tiles_c, chunked;
    int64_t s_r, s_c, ld;        // ld: destination row pitch in elements (>= ntaps * cpad)
    int32_t tapoff[MAXT];
    int32_t blk0, nblk;
};

__global__ void __launch_bounds__(256) wprep_batch_kernel(const WJob* __restrict__ jobs, const int* __restrict__ blk2job) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[RB * CB * MAXT];
    const WJob& J = jobs[blk2job[blockIdx.x]];      // every thread reads the (L1-broadcast) job record directly
    const int local = blockIdx.x - J.blk0;
    const int r0 = (local / J.tiles_c) * RB, c0 = (local % J.tiles_c) * CB;
    const int ts = J.ts;
    const bool pow2 = (ts == 16);                                        // 4x4 kernels: XOR-swizzle the tap index (bank conflicts)
    const unsigned magic = (65536u + ts - 1) / ts;                        // j / ts for j < 4096 without an integer division
    const int nrt = min(RB, J.nr - r0), nct = min(CB, J.nc - c0);       // valid rows / cols of this tile (may be <= 0)
    // smem element (rr, cc, tt) lives at (rr*CB + cc)*ts + (pow2 ? tt ^ (cc & 15) : tt)
    // The kernel is a pure HBM stream (fp32 in, 16 bit out): all loads of a thread are issued before the first shared-memory
    // store (8 independent 4-byte loads in flight per thread) -- with one load in flight it ran at 1/5 of the copy bandwidth.
    if (nrt > 0 && nct > 0) {
        if (J.s_c < J.s_r) {
            // columns are the inner dimension: for each row, [nct * ts] floats are contiguous -> straight copy
            const int run = nct * ts;
            const float* sp0 = J.src + (int64_t)r0 * J.s_r + (int64_t)c0 * J.s_c;
            for (int j = threadIdx.x; j < run; j += 256) {
                float v[RB];
#pragma unroll
                for (int rr = 0; rr < RB; rr++) v[rr] = (rr < nrt) ? __ldcs(sp0 + (int64_t)rr * J.s_r + j) : 0.f;
                int d = j;
                if (pow2) { const int cc = j >> 4; d = (cc << 4) | ((j & 15) ^ (cc & 15)); }
#pragma unroll
                for (int rr = 0; rr < RB; rr++) tile[rr * CB * ts + d] = v[rr];
            }
        } else {
            // rows are the inner dimension: for each column, [nrt * ts] (<= 128) floats are contiguous; a warp owns 8 columns
            const int run = nrt * ts;
            const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
            const float* sp0 = J.src + (int64_t)c0 * J.s_c + (int64_t)r0 * J.s_r;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                float v[4][4];
#pragma unroll
                for (int ci = 0; ci < 4; ci++) {
                    const int cc = wq + 8 * (half * 4 + ci);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = lane + 32 * u;
                        v[ci][u] = (cc < nct && j < run) ? __ldcs(sp0 + (int64_t)cc * J.s_c + j) : 0.f;
                    }
                }
#pragma unroll
                for (int ci = 0; ci < 4; ci++) {
                    const int cc = wq + 8 * (half * 4 + ci);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = lane + 32 * u;
                        if (cc < nct && j < run) {
                            const int rr = pow2 ? (j >> 4) : (int)(((unsigned)j * magic) >> 16);
                            int tt = j - rr * ts;
                            if (pow2) tt ^= (cc & 15);
                            tile[(rr * CB + cc) * ts + tt] = v[ci][u];
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int ncw = min(CB, J.cpad - c0);
    const int nrw = min(RB, J.nr_pad - r0);
    unsigned short* dst = (unsigned short*)J.dst;
    // thread = (8-column group cg, tap lane tl, row lane rl): one 16-byte store per (row, tap, column group)
    const int cg = threadIdx.x & 7, tl = (threadIdx.x >> 3) & 15, rl = threadIdx.x >> 7;     // 8 x 16 x 2
    if (cg * 8 < ncw && tl < J.ntaps) {
        const int tsel = J.tapoff[tl];
        for (int rr = rl; rr < nrw; rr += 2) {
            __align__(16) unsigned short o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int cc = cg * 8 + e;
                float v = 0.f;
                if (rr < nrt && cc < nct) v = tile[(rr * CB + cc) * ts + (pow2 ? (tsel ^ (cc & 15)) : tsel)];
                o[e] = f2b16(v, J.dtype);
            }
            const int c = c0 + cg * 8;
            const int64_t col = J.chunked ? ((int64_t)((c >> 6) * J.ntaps + tl) * 64 + (c & 63)) : ((int64_t)tl * J.cpad + c);
            *(uint4*)(dst + (int64_t)(r0 + rr) * J.ld + col) = *(const uint4*)o;
        }
    }
}

}  // namespace

struct rnr_wprep_plan {
    WJob* d_jobs = nullptr;
    int* d_blk2job = nullptr;
    int njobs = 0;
    int nblocks = 0;
};

extern "C" int rnr_wprep_plan_create(const rnr_wprep_job_t* jobs, int njobs, rnr_wprep_plan_t** out) {
    RNR_REQUIRE(jobs && out && njobs >= 1, "rnr_wprep_plan_create: bad arguments");
    std::vector<WJob> h(njobs);
    int blk = 0;
    for (int i = 0; i < njobs; i++) {
        const rnr_wprep_job_t& s = jobs[i];
        RNR_REQUIRE(s.dst_dtype == RNR_F16 || s.dst_dtype == RNR_BF16, "weight prep: destination must be 16-bit");
        RNR_REQUIRE(s.ntaps >= 1 && s.ntaps <= MAXT, "weight prep: 1..%d taps, got %d", MAXT, s.ntaps);
        const int64_t ts = s.s_r < s.s_c ? s.s_r : s.s_c;
        RNR_REQUIRE(ts >= 1 && ts <= MAXT, "weight prep: inner tap stride %lld out of range", (long long)ts);
        WJob& d = h[i];
        d.src = s.src; d.dst = s.dst; d.dtype = s.dst_dtype;
        d.nr = s.nr; d.nr_pad = s.nr_pad; d.nc = s.nc; d.cpad = s.cpad; d.ntaps = s.ntaps; d.ts = (int)ts;
        d.s_r = s.s_r; d.s_c = s.s_c; d.chunked = s.chunked;
        d.ld = s.ld > 0 ? s.ld : (int64_t)s.ntaps * s.cpad;
        RNR_REQUIRE(d.ld >= (int64_t)s.ntaps * s.cpad && d.ld % 8 == 0, "weight prep: row pitch %lld too small / not a multiple of 8", (long long)d.ld);
        RNR_REQUIRE(s.cpad % 8 == 0 && ((uintptr_t)s.dst & 15) == 0, "weight prep: cpad %% 8 == 0 and a 16-byte aligned destination required (cpad %d)", s.cpad);
        RNR_REQUIRE(!s.chunked || s.cpad % 64 == 0, "weight prep: chunk-major layout needs cpad %% 64 == 0 (got %d)", s.cpad);
        for (int t = 0; t < s.ntaps; t++) {
            RNR_REQUIRE(s.tapoff[t] >= 0 && s.tapoff[t] < ts, "weight prep: tap offset %d outside [0,%lld)", s.tapoff[t], (long long)ts);
            d.tapoff[t] = s.tapoff[t];
        }
        d.tiles_c = rnr_cdiv(s.cpad, CB);
        d.blk0 = blk;
        d.nblk = rnr_cdiv(s.nr_pad, RB) * d.tiles_c;
        blk += d.nblk;
    }
    rnr_wprep_plan* p = new rnr_wprep_plan();
    p->njobs = njobs;
    p->nblocks = blk;
    RNR_CHECK(cudaMalloc(&p->d_jobs, sizeof(WJob) * njobs));
    RNR_CHECK(cudaMemcpy(p->d_jobs, h.data(), sizeof(WJob) * njobs, cudaMemcpyHostToDevice));
    std::vector<int> b2j(blk);
    for (int i = 0; i < njobs; i++)
        for (int b = 0; b < h[i].nblk; b++) b2j[h[i].blk0 + b] = i;
    RNR_CHECK(cudaMalloc(&p->d_blk2job, sizeof(int) * blk));
    RNR_CHECK(cudaMemcpy(p->d_blk2job, b2j.data(), sizeof(int) * blk, cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}

extern "C" void rnr_wprep_plan_destroy(rnr_wprep_plan_t* p) {
    if (!p) return;
    cudaFree(p->d_jobs);
    cudaFree(p->d_blk2job);
    delete p;
}

extern "C" int rnr_wprep_run(const rnr_wprep_plan_t* p, void* stream) {
    RNR_REQUIRE(p, "rnr_wprep_run: null plan");
    RNR_PDL_LAUNCH(wprep_batch_kernel, p->nblocks, 256, 0, stream, p->d_jobs, p->d_blk2job);
    RNR_LAUNCH_CHECK();
    return 0;
}

// -------------------------------------------------------------------------------------------------------------------
// Weight-gradient un-transpose.  The tcgen05 weight-gradient kernel accumulates dW in GEMM order [tap][co][ci] (ci contiguous:
// every thread of its epilogue owns one co row and issues 128-bit vector reductions) instead of the parameter's own order
// ([co][ci][kh][kw] for nn.Conv2d, [ci][co][kh][kw] for nn.ConvTranspose2d), where the 4-byte reductions of a warp land in 32
// different sectors.  One batched launch per step moves every layer's gradient into the flat fp32 gradient buffer that
// torch.optim.Adam reads:   dst[co*s_co + ci*s_ci + t] = src[(t*cout + co)*cin + ci].
// -------------------------------------------------------------------------------------------------------------------
namespace {

struct WUnJob {
    const float* src;
    float* dst;
    int32_t cout, cin, ntaps, tiles_c;
    int64_t s_co, s_ci;
    int32_t blk0, nblk;
};

constexpr int UCO = 4;      // co rows per block (one per group of 64 threads)

__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const WUnJob* __restrict__ jobs, const int* __restrict__ blk2job) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[UCO][MAXT][CB + 1];
    const WUnJob& J = jobs[blk2job[blockIdx.x]];
    const int local = blockIdx.x - J.blk0;
    const int cog = local / J.tiles_c, c0 = (local - cog * J.tiles_c) * CB;
    const int r = threadIdx.x >> 6, cc = threadIdx.x & 63;     // thread = (co row of the block, input channel of the chunk)
    const int co = cog * UCO + r;
    const int nc = min(CB, J.cin - c0);
    const int ntaps = J.ntaps;
    // loads: every thread fetches its channel of all taps (ntaps independent 4-byte loads in flight; a warp reads 128 B rows)
    if (co < J.cout && cc < nc) {
        const float* sp = J.src + (int64_t)co * J.cin + c0 + cc;
        const int64_t tstride = (int64_t)J.cout * J.cin;
        float v[MAXT];
#pragma unroll
        for (int t = 0; t < MAXT; t++) v[t] = (t < ntaps) ? __ldcs(sp + t * tstride) : 0.f;
#pragma unroll
        for (int t = 0; t < MAXT; t++)
            if (t < ntaps) tile[r][t][cc] = v[t];
    }
    __syncthreads();
    // stores: row r of the block is a contiguous run of nc * ntaps floats when s_ci == ntaps (nn.Conv2d); j / ntaps by a
    // multiply-shift (j < 1024, ntaps <= 16: exact)
    if (co < J.cout) {
        const unsigned magic = (65536u + ntaps - 1) / ntaps;
        float* dp = J.dst + (int64_t)co * J.s_co + (int64_t)c0 * J.s_ci;
        const int run = nc * ntaps;
        for (int j = cc; j < run; j += 64) {
            const int ci = (int)(((unsigned)j * magic) >> 16), t = j - ci * ntaps;
            dp[(int64_t)ci * J.s_ci + t] = tile[r][t][ci];
        }
    }
}

}  // namespace

struct rnr_wunpack_plan {
    WUnJob* d_jobs = nullptr;
    int* d_blk2job = nullptr;
    int nblocks = 0;
};

extern "C" int rnr_wgrad_unpack_plan_create(const rnr_wunpack_job_t* jobs, int njobs, rnr_wunpack_plan_t** out) {
    RNR_REQUIRE(jobs && out && njobs >= 1, "rnr_wgrad_unpack_plan_create: bad arguments");
    std::vector<WUnJob> h(njobs);
    int blk = 0;
    for (int i = 0; i < njobs; i++) {
        const rnr_wunpack_job_t& s = jobs[i];
        RNR_REQUIRE(s.ntaps >= 1 && s.ntaps <= MAXT, "wgrad unpack: 1..%d taps, got %d", MAXT, s.ntaps);
        WUnJob& d = h[i];
        d.src = s.src; d.dst = s.dst; d.cout = s.cout; d.cin = s.cin; d.ntaps = s.ntaps; d.s_co = s.s_co; d.s_ci = s.s_ci;
        d.tiles_c = rnr_cdiv(s.cin, CB);
        d.blk0 = blk;
        d.nblk = rnr_cdiv(s.cout, UCO) * d.tiles_c;
        blk += d.nblk;
    }
    rnr_wunpack_plan* p = new rnr_wunpack_plan();
    p->nblocks = blk;
    RNR_CHECK(cudaMalloc(&p->d_jobs, sizeof(WUnJob) * njobs));
    RNR_CHECK(cudaMemcpy(p->d_jobs, h.data(), sizeof(WUnJob) * njobs, cudaMemcpyHostToDevice));
    std::vector<int> b2j(blk);
    for (int i = 0; i < njobs; i++)
        for (int b = 0; b < h[i].nblk; b++) b2j[h[i].blk0 + b] = i;
    RNR_CHECK(cudaMalloc(&p->d_blk2job, sizeof(int) * blk));
    RNR_CHECK(cudaMemcpy(p->d_blk2job, b2j.data(), sizeof(int) * blk, cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}

extern "C" void rnr_wgrad_unpack_plan_destroy(rnr_wunpack_plan_t* p) {
    if (!p) return;
    cudaFree(p->d_jobs);
    cudaFree(p->d_blk2job);
    delete p;
}

extern "C" int rnr_wgrad_unpack_run(const rnr_wunpack_plan_t* p, void* stream) {
    RNR_REQUIRE(p, "rnr_wgrad_unpack_run: null plan");
    RNR_PDL_LAUNCH(wgrad_unpack_kernel, p->nblocks, 256, 0, stream, p->d_jobs, p->d_blk2job);
    RNR_LAUNCH_CHECK();
    return 0;
}

