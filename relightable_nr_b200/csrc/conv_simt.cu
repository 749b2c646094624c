// Generic implicit-GEMM convolution / weight-gradient problems: plan management and the
// SIMT (CUDA-core) validation kernels.  The tcgen05 kernels live in conv_tc.cu and consume
// exactly the same problem description, so this file is the on-GPU cross-check for them.
//
// Semantics restated from the reference's use of nn.Conv2d / nn.ConvTranspose2d in
// pytorch_prototyping/pytorch_prototyping.py:112-115 (Conv2dSame), :155-160 (UpBlock
// ConvTranspose2d 4x4 s2 p1), :242-264 (DownBlock 3x3 s1 + 4x4 s2 after ReflectionPad2d(1)).
#include "conv_internal.cuh"
#include <stdlib.h>
#include <vector>

// ---------------------------------------------------------------------------------------------
// SIMT conv: 128 (pixels) x 64 (channels) tile per CTA, 256 threads, 8x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int SBM = 128, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];
    __shared__ float s_sum[SBN], s_sq[SBN];

    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x;
    const int n0 = blockIdx.y * SBN;
    const int tx_ = tile_m % p.tiles_x;
    const int ty_ = (tile_m / p.tiles_x) % p.tiles_y;
    const int n_ = tile_m / (p.tiles_x * p.tiles_y);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    // A loader: row = tid/2, 8 channels starting at (tid%2)*8
    const int a_row = tid >> 1, a_k0 = (tid & 1) * 8;
    const int a_y = ty_ * p.th + a_row / p.tw;
    const int a_x = tx_ * p.tw + a_row % p.tw;
    // B loader: col = tid/4, 4 k starting at (tid%4)*4
    const int b_col = tid >> 2, b_k0 = (tid & 3) * 4;
    const int rg = tid >> 4, cg = tid & 15;

    for (int j = 0; j < p.n_ksteps; j++) {
        const rnr_kstep_t ks = p.ksteps[j];
        const ViewD& v = p.views[ks.view];
        const int x = a_x + ks.dx, y = a_y + ks.dy;
        const bool inb = (x >= 0 && x < v.dim[1] && y >= 0 && y < v.dim[2] && n_ < v.dim[3]);
        const unsigned short* abase = (const unsigned short*)v.ptr + x * v.stride[1] + y * v.stride[2] + n_ * v.stride[3];
        for (int sub = 0; sub < p.bk; sub += SBK) {
            // ---- load A ----
            {
                const int c = ks.c0 + sub + a_k0;
                float vals[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    float f = 0.f;
                    if (inb && (c + e) < v.dim[0]) f = cvt16(abase[c + e], p.ab_dtype);
                    vals[e] = f;
                }
#pragma unroll
                for (int e = 0; e < 8; e++) As[a_k0 + e][a_row] = vals[e];
            }
            // ---- load B ----
            {
                const int row = n0 + b_col;
                const unsigned short* wb = (const unsigned short*)p.wmat + (int64_t)row * p.ldw + (int64_t)j * p.bk + sub + b_k0;
#pragma unroll
                for (int e = 0; e < 4; e++) Bs[b_k0 + e][b_col] = (row < p.n_rows_w) ? cvt16(wb[e], p.ab_dtype) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SBK; k++) {
                float a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = As[k][rg * 8 + i];
#pragma unroll
                for (int jj = 0; jj < 4; jj++) b[jj] = Bs[k][cg * 4 + jj];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) acc[i][jj] = fmaf(a[i], b[jj], acc[i][jj]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue ----
    if (p.epi & RNR_EPI_STATS) {
        if (tid < SBN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
        __syncthreads();
    }
    float csum[4] = {0, 0, 0, 0}, csq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = rg * 8 + i;
        const int y = ty_ * p.th + r / p.tw, x = tx_ * p.tw + r % p.tw;
        const bool valid = (y < p.mY && x < p.mX);
        const int64_t obase = (int64_t)n_ * p.out_sn + (int64_t)(y * p.out_my + p.out_py) * p.out_sy +
                              (int64_t)(x * p.out_mx + p.out_px) * p.out_sx;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int co = n0 + cg * 4 + jj;
            if (!valid || co >= p.cout) continue;
            float vv = acc[i][jj];
            if (p.epi & RNR_EPI_BIAS) vv += p.bias[co];
            if (p.epi & RNR_EPI_TANH) vv = tanhf(vv);
            st_out(p.out, obase + co, vv, p.out_dtype);
            csum[jj] += vv;
            csq[jj] += vv * vv;
        }
    }
    if (p.epi & RNR_EPI_STATS) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            atomicAdd(&s_sum[cg * 4 + jj], csum[jj]);
            atomicAdd(&s_sq[cg * 4 + jj], csq[jj]);
        }
        __syncthreads();
        if (tid < SBN && n0 + tid < p.cout) {
            p.stats[((int64_t)tile_m * 2 + 0) * p.ldstats + n0 + tid] = s_sum[tid];
            p.stats[((int64_t)tile_m * 2 + 1) * p.ldstats + n0 + tid] = s_sq[tid];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT wgrad: 64 (co) x 64 (ci) output tile per CTA for one tap, split over the pixel space.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradParams p, const int* __restrict__ tile_tab,
                                                       int splitk) {
    __shared__ float Gs[16][64 + 4];
    __shared__ float As[16][64 + 4];
    const int tid = threadIdx.x;
    const int* tt = tile_tab + blockIdx.x * 3;
    const rnr_wtap_t tap = p.taps[tt[0]];
    const int ci_b = tt[1] * 64, co_b = tt[2] * 64;
    const ViewD& av = p.aviews[tap.view];
    const ViewD& gv = p.gviews[tap.gview];

    const int64_t npix = (int64_t)p.mN * p.mY * p.mX;
    const int64_t per = (npix + splitk - 1) / splitk;
    const int64_t pb = (int64_t)blockIdx.y * per;
    const int64_t pe = (pb + per < npix) ? pb + per : npix;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    const int l_pix = tid >> 4;          // 0..15
    const int l_c = (tid & 15) * 4;      // 4 channels
    const int rg = tid >> 4, cg = tid & 15;

    for (int64_t p0 = pb; p0 < pe; p0 += 16) {
        const int64_t pix = p0 + l_pix;
        float g4[4] = {0, 0, 0, 0}, a4[4] = {0, 0, 0, 0};
        if (pix < pe) {
            const int x = (int)(pix % p.mX);
            const int y = (int)((pix / p.mX) % p.mY);
            const int n = (int)(pix / ((int64_t)p.mX * p.mY));
            if (x < gv.dim[1] && y < gv.dim[2] && n < gv.dim[3]) {
                const unsigned short* gp = (const unsigned short*)gv.ptr + x * gv.stride[1] + y * gv.stride[2] + n * gv.stride[3];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int co = co_b + l_c + e;
                    if (co < p.cout && co < gv.dim[0]) g4[e] = cvt16(gp[co], p.g_dtype);
                }
            }
            const int ax = x + tap.dx, ay = y + tap.dy;
            if (ax >= 0 && ax < av.dim[1] && ay >= 0 && ay < av.dim[2] && n < av.dim[3]) {
                const unsigned short* ap = (const unsigned short*)av.ptr + ax * av.stride[1] + ay * av.stride[2] + n * av.stride[3];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int ci = ci_b + l_c + e;
                    if (ci < tap.nci && (tap.c0 + ci) < av.dim[0]) a4[e] = cvt16(ap[tap.c0 + ci], p.a_dtype);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) { Gs[l_pix][l_c + e] = g4[e]; As[l_pix][l_c + e] = a4[e]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float g[4], a[4];
#pragma unroll
            for (int i = 0; i < 4; i++) g[i] = Gs[k][rg * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) a[j] = As[k][cg * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int co = co_b + rg * 4 + i;
        if (co >= p.cout) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int ci = ci_b + cg * 4 + j;
            if (ci >= tap.nci) continue;
            atomicAdd(p.dw + co * p.s_co + (int64_t)(tap.ci0 + ci) * p.s_ci + tap.off, acc[i][j]);
        }
    }
}

__global__ void weight_prep_kernel(const float* __restrict__ src, void* __restrict__ dst, int dst_dtype,
                                   int nr, int nr_pad, int nc, int cpad, int ntaps,
                                   int64_t s_r, int64_t s_c, const int32_t* __restrict__ tapoff, int chunked) {
    const int64_t total = (int64_t)nr_pad * ntaps * cpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpad);
        const int t = (int)((i / cpad) % ntaps);
        const int r = (int)(i / ((int64_t)cpad * ntaps));
        float v = 0.f;
        if (r < nr && c < nc) v = src[r * s_r + c * s_c + tapoff[t]];
        const int64_t col = chunked ? ((int64_t)((c >> 6) * ntaps + t) * 64 + (c & 63)) : ((int64_t)t * cpad + c);
        ((unsigned short*)dst)[(int64_t)r * ntaps * cpad + col] = f2b16(v, dst_dtype);
    }
}

void copy_view(ViewD& d, const rnr_view_t& s) {
    d.ptr = s.ptr;
    for (int i = 0; i < 4; i++) { d.dim[i] = s.dim[i]; d.stride[i] = s.stride[i]; }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int conv_plan_create(const rnr_conv_problem_t* prob, int nsub, int impl, rnr_conv_plan_t** out);

extern "C" int rnr_conv_plan_create(const rnr_conv_problem_t* prob, int impl, rnr_conv_plan_t** out) {
    return conv_plan_create(prob, 1, impl, out);
}

/* `n` sub-problems that differ only in tap offsets, weight matrix (stacked row-wise in one buffer) and output parity, run as ONE
 * launch of the tcgen05 halo kernel.  Returns cudaErrorNotSupported (and no plan) when the problems cannot be fused -- the caller
 * then creates one plan per problem. */
extern "C" int rnr_conv_plan_create_multi(const rnr_conv_problem_t* probs, int n, int impl, rnr_conv_plan_t** out) {
    RNR_REQUIRE(probs && out && n >= 1 && n <= 4, "rnr_conv_plan_create_multi: 1..4 sub-problems");
    if (n > 1 && impl != 1) { rnr_set_error("rnr_conv_plan_create_multi: only the tcgen05 implementation fuses sub-problems"); return (int)cudaErrorNotSupported; }
    return conv_plan_create(probs, n, impl, out);
}

static int conv_plan_create(const rnr_conv_problem_t* prob, int nsub, int impl, rnr_conv_plan_t** out) {
    RNR_REQUIRE(prob && out, "rnr_conv_plan_create: null argument");
    RNR_REQUIRE(prob->th * prob->tw == 128, "rnr_conv_plan_create: tile must hold 128 pixels (th=%d tw=%d)", prob->th, prob->tw);
    RNR_REQUIRE(prob->bk == 16 || prob->bk == 32 || prob->bk == 64, "rnr_conv_plan_create: bk must be 16/32/64");
    RNR_REQUIRE(prob->n_views >= 1 && prob->n_views <= RNR_MAX_VIEWS, "rnr_conv_plan_create: bad n_views");
    RNR_REQUIRE(prob->ab_dtype == RNR_F16 || prob->ab_dtype == RNR_BF16, "rnr_conv_plan_create: operands must be 16-bit");
    rnr_conv_plan* pl = new rnr_conv_plan();
    memset(pl, 0, sizeof(*pl));
    ConvParams& p = pl->p;
    for (int i = 0; i < prob->n_views; i++) copy_view(p.views[i], prob->views[i]);
    p.n_ksteps = prob->n_ksteps; p.bk = prob->bk; p.ab_dtype = prob->ab_dtype;
    p.wmat = prob->wmat; p.ldw = prob->n_ksteps * prob->bk; p.n_rows_w = prob->n_rows_w; p.cout = prob->cout;
    p.mN = prob->mN; p.mY = prob->mY; p.mX = prob->mX; p.th = prob->th; p.tw = prob->tw;
    p.tiles_y = rnr_cdiv(prob->mY, prob->th); p.tiles_x = rnr_cdiv(prob->mX, prob->tw);
    p.tiles_m = p.tiles_y * p.tiles_x * prob->mN;
    p.out = prob->out; p.out_dtype = prob->out_dtype;
    p.out_sn = prob->out_sn; p.out_sy = prob->out_sy; p.out_sx = prob->out_sx;
    p.out_my = prob->out_my; p.out_mx = prob->out_mx; p.out_py = prob->out_py; p.out_px = prob->out_px;
    p.epi = prob->epi; p.bias = prob->bias; p.stats = prob->stats; p.ldstats = prob->ldstats;
    pl->impl = impl;
    cudaError_t e = cudaMalloc(&pl->d_ksteps, sizeof(rnr_kstep_t) * prob->n_ksteps);
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_ksteps, prob->ksteps, sizeof(rnr_kstep_t) * prob->n_ksteps, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rnr_set_error("rnr_conv_plan_create: %s", cudaGetErrorString(e)); delete pl; return (int)e; }
    p.ksteps = pl->d_ksteps;
    if (impl == 1) {
        int rc = rnr_conv_halo_prepare(pl, prob, nsub);      // shared-memory halo reuse when the problem fits ...
        if (rc == 0 && !pl->halo && nsub > 1) {
            rnr_set_error("rnr_conv_plan_create_multi: sub-problems cannot be fused by the halo kernel");
            rc = (int)cudaErrorNotSupported;
        }
        if (rc == 0 && !pl->halo) rc = rnr_conv_tc_prepare(pl, prob);   // ... else one A tile per tap
        if (rc != 0) { cudaFree(pl->d_ksteps); if (pl->d_groups) cudaFree(pl->d_groups); if (pl->d_taps) cudaFree(pl->d_taps); delete pl; return rc; }
    }
    *out = pl;
    return 0;
}

extern "C" void rnr_conv_plan_destroy(rnr_conv_plan_t* plan) {
    if (!plan) return;
    cudaFree(plan->d_ksteps);
    if (plan->d_groups) cudaFree(plan->d_groups);
    if (plan->d_taps) cudaFree(plan->d_taps);
    delete plan;
}

/* rows of the BatchNorm partial-sum buffer this plan writes: one per M tile (SIMT / per-tap kernel) or one per CTA (halo kernel) */
extern "C" int rnr_conv_plan_stat_rows(const rnr_conv_plan_t* plan) {
    if (!plan) return 0;
    return (plan->impl == 1 && plan->halo) ? plan->grid : plan->p.tiles_m;
}

extern "C" int rnr_conv_plan_tiles_m(const rnr_conv_plan_t* plan) { return plan ? plan->p.tiles_m : 0; }

extern "C" int rnr_conv_run(const rnr_conv_plan_t* plan, void* stream) {
    RNR_REQUIRE(plan, "rnr_conv_run: null plan");
    if (plan->impl == 1) return plan->halo ? rnr_conv_halo_run(plan, (cudaStream_t)stream) : rnr_conv_tc_run(plan, (cudaStream_t)stream);
    dim3 grid(plan->p.tiles_m, rnr_cdiv(plan->p.cout, SBN));
    conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan->p);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_wgrad_plan_create(const rnr_wgrad_problem_t* prob, int impl, rnr_wgrad_plan_t** out) {
    RNR_REQUIRE(prob && out, "rnr_wgrad_plan_create: null argument");
    RNR_REQUIRE(prob->n_aviews >= 1 && prob->n_aviews <= RNR_MAX_VIEWS && prob->n_gviews >= 1 && prob->n_gviews <= 4,
                "rnr_wgrad_plan_create: bad view counts");
    rnr_wgrad_plan* pl = new rnr_wgrad_plan();
    memset(pl, 0, sizeof(*pl));
    WgradParams& p = pl->p;
    for (int i = 0; i < prob->n_aviews; i++) copy_view(p.aviews[i], prob->aviews[i]);
    for (int i = 0; i < prob->n_gviews; i++) copy_view(p.gviews[i], prob->gviews[i]);
    p.n_taps = prob->n_taps; p.a_dtype = prob->a_dtype; p.g_dtype = prob->g_dtype; p.cout = prob->cout;
    p.mN = prob->mN; p.mY = prob->mY; p.mX = prob->mX; p.dw = prob->dw; p.s_co = prob->s_co; p.s_ci = prob->s_ci;
    pl->impl = impl;
    pl->h_taps = (rnr_wtap_t*)malloc(sizeof(rnr_wtap_t) * prob->n_taps);
    memcpy(pl->h_taps, prob->taps, sizeof(rnr_wtap_t) * prob->n_taps);
    cudaError_t e = cudaMalloc(&pl->d_taps, sizeof(rnr_wtap_t) * prob->n_taps);
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_taps, prob->taps, sizeof(rnr_wtap_t) * prob->n_taps, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rnr_set_error("rnr_wgrad_plan_create: %s", cudaGetErrorString(e)); free(pl->h_taps); delete pl; return (int)e; }
    p.taps = pl->d_taps;
    if (impl == 1) {
        int rc = rnr_wgrad_halo_prepare(pl, prob);           // halo reuse + one accumulator per tap where the layer is large ...
        if (rc == 0 && !pl->halo) rc = rnr_wgrad_tc_prepare(pl, prob);   // ... else one box pair per tap
        if (rc != 0) { cudaFree(pl->d_taps); free(pl->h_taps); delete pl; return rc; }
    } else {
        std::vector<int> tab;
        for (int t = 0; t < prob->n_taps; t++)
            for (int cb = 0; cb < rnr_cdiv(prob->taps[t].nci, 64); cb++)
                for (int ob = 0; ob < rnr_cdiv(prob->cout, 64); ob++) { tab.push_back(t); tab.push_back(cb); tab.push_back(ob); }
        pl->n_tiles = (int)tab.size() / 3;
        e = cudaMalloc(&pl->d_tile_tab, tab.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(pl->d_tile_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { rnr_set_error("rnr_wgrad_plan_create: %s", cudaGetErrorString(e)); cudaFree(pl->d_taps); free(pl->h_taps); delete pl; return (int)e; }
        const long long npix = (long long)prob->mN * prob->mY * prob->mX;
        int sk = (int)((148 * 8 + pl->n_tiles - 1) / pl->n_tiles);
        long long maxsk = (npix + 255) / 256;
        if (sk > maxsk) sk = (int)maxsk;
        if (sk < 1) sk = 1;
        if (sk > 65535) sk = 65535;
        pl->splitk = sk;
    }
    *out = pl;
    return 0;
}

extern "C" void rnr_wgrad_plan_destroy(rnr_wgrad_plan_t* plan) {
    if (!plan) return;
    cudaFree(plan->d_taps);
    if (plan->d_tile_tab) cudaFree(plan->d_tile_tab);
    if (plan->d_work_tab) cudaFree(plan->d_work_tab);
    free(plan->h_taps);
    delete plan;
}

extern "C" int rnr_wgrad_run(const rnr_wgrad_plan_t* plan, void* stream) {
    RNR_REQUIRE(plan, "rnr_wgrad_run: null plan");
    if (plan->impl == 1) return plan->halo ? rnr_wgrad_halo_run(plan, (cudaStream_t)stream) : rnr_wgrad_tc_run(plan, (cudaStream_t)stream);
    dim3 grid(plan->n_tiles, plan->splitk);
    wgrad_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan->p, plan->d_tile_tab, plan->splitk);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_weight_prep(const float* src, void* dst, int dst_dtype, int nr, int nr_pad, int nc, int cpad,
                               int ntaps, int64_t s_r, int64_t s_c, const int32_t* tapoff_dev, int chunked, void* stream) {
    RNR_REQUIRE(!chunked || cpad % 64 == 0, "rnr_weight_prep: chunk-major layout needs cpad %% 64 == 0");
    RNR_REQUIRE(dst_dtype == RNR_F16 || dst_dtype == RNR_BF16, "rnr_weight_prep: dst must be 16-bit");
    const int64_t total = (int64_t)nr_pad * ntaps * cpad;
    int blocks = rnr_cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    weight_prep_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, dst_dtype, nr, nr_pad, nc, cpad, ntaps, s_r, s_c, tapoff_dev, chunked);
    RNR_LAUNCH_CHECK();
    return 0;
}
