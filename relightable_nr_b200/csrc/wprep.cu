// Batched weight preparation: fp32 master weights -> the 16-bit K-major GEMM matrices of every layer, ONE launch.
//
// The U-Net has 22 live layers; each needs a forward matrix (4 for a ConvTranspose: one per output parity) and a
// data-gradient matrix (4 for a stride-2 conv) re-derived from the fp32 parameters after every optimiser step:
// 74 small strided transposes.  A plan holds the job table in device memory; one persistent-style launch walks
// (job, 8-row x 64-column tile) work items.  Each tile is staged through shared memory so that both the fp32 source
// (contiguous along whichever of row/column is the inner dimension of the parameter tensor) and the 16-bit
// destination (contiguous along the column = input-channel dimension) move in full sectors.
//   dst[r*ntaps*cpad + t*cpad + c] = src[r*s_r + c*s_c + tapoff[t]]     (0 for r >= nr or c >= nc)
// Replaces the per-layer weight reshapes that cuDNN does inside nn.Conv2d / nn.ConvTranspose2d
// (pytorch_prototyping/pytorch_prototyping.py:112-115,155-160,242-264).
#include "common.cuh"
#include <vector>

namespace {

constexpr int RB = 8, CB = 64, MAXT = 16;

struct WJob {
    const float* src;
    void* dst;
    int32_t dtype, nr, nr_pad, nc, cpad, ntaps, ts, tiles_c;
    int64_t s_r, s_c;
    int32_t tapoff[MAXT];
    int32_t blk0, nblk;
};

__global__ void __launch_bounds__(256) wprep_batch_kernel(const WJob* __restrict__ jobs, int njobs) {
    __shared__ float tile[RB * CB * (MAXT + 1)];
    __shared__ WJob J;
    // locate the job of this block (block offsets are ascending)
    if (threadIdx.x == 0) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].blk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        J = jobs[lo];
    }
    __syncthreads();
    const int local = blockIdx.x - J.blk0;
    const int r0 = (local / J.tiles_c) * RB, c0 = (local % J.tiles_c) * CB;
    const int ts = J.ts, tsp = ts + 1;
    const int nrt = min(RB, J.nr - r0), nct = min(CB, J.nc - c0);       // valid rows / cols of this tile (may be <= 0)
    if (nrt > 0 && nct > 0) {
        if (J.s_c < J.s_r) {
            // columns are the inner dimension: for each row, [nct * ts] floats are contiguous
            const int run = nct * ts;
            for (int i = threadIdx.x; i < nrt * run; i += 256) {
                const int rr = i / run, j = i - rr * run;
                const int cc = j / ts, tt = j - cc * ts;
                tile[(rr * CB + cc) * tsp + tt] = J.src[(int64_t)(r0 + rr) * J.s_r + (int64_t)c0 * J.s_c + j];
            }
        } else {
            // rows are the inner dimension: for each column, [nrt * ts] floats are contiguous
            const int run = nrt * ts;
            for (int i = threadIdx.x; i < nct * run; i += 256) {
                const int cc = i / run, j = i - cc * run;
                const int rr = j / ts, tt = j - rr * ts;
                tile[(rr * CB + cc) * tsp + tt] = J.src[(int64_t)(c0 + cc) * J.s_c + (int64_t)r0 * J.s_r + j];
            }
        }
    }
    __syncthreads();
    const int ncw = min(CB, J.cpad - c0);
    const int nrw = min(RB, J.nr_pad - r0);
    unsigned short* dst = (unsigned short*)J.dst;
    for (int i = threadIdx.x; i < nrw * J.ntaps * CB; i += 256) {
        const int cc = i % CB;
        const int rt = i / CB;
        const int t = rt % J.ntaps, rr = rt / J.ntaps;
        if (cc >= ncw) continue;
        float v = 0.f;
        if (rr < nrt && cc < nct) v = tile[(rr * CB + cc) * tsp + J.tapoff[t]];
        dst[((int64_t)(r0 + rr) * J.ntaps + t) * J.cpad + c0 + cc] = f2b16(v, J.dtype);
    }
}

}  // namespace

struct rnr_wprep_plan {
    WJob* d_jobs = nullptr;
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
        d.s_r = s.s_r; d.s_c = s.s_c;
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
    *out = p;
    return 0;
}

extern "C" void rnr_wprep_plan_destroy(rnr_wprep_plan_t* p) {
    if (!p) return;
    cudaFree(p->d_jobs);
    delete p;
}

extern "C" int rnr_wprep_run(const rnr_wprep_plan_t* p, void* stream) {
    RNR_REQUIRE(p, "rnr_wprep_run: null plan");
    wprep_batch_kernel<<<p->nblocks, 256, 0, (cudaStream_t)stream>>>(p->d_jobs, p->njobs);
    RNR_LAUNCH_CHECK();
    return 0;
}
