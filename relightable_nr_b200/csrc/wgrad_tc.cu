// tcgen05 weight-gradient kernel.
//
//   dW[co, ci0+ci] (+tap offset) += sum over pixels  G[pix, co] * A[pix + tap, c0 + ci]
//
// As a GEMM:  D[M = co (128), N = ci (<=128)]  +=  G^T [M x K]  *  A [K x N],   K = pixels.
// Both operands are "MN-major": the contiguous memory dimension (channels) is the M resp. N dimension,
// so the tiles are used exactly as TMA writes them ([pixel rows x 64 channels], 128-byte rows, SWIZZLE_128B)
// with MN-major UMMA descriptors -- no transposition pass.  Both operands must have the SAME 16-bit format: the
// instruction descriptor has separate A / B format fields, but kind::f16 with fp16 x bf16 raises an illegal-instruction
// fault on sm_100a (measured), which is why the engine keeps a bf16 copy of the activations for this kernel.
//
// Work item = (tap entry, co block of 128, ci chunk of <=128, pixel range); one persistent CTA per SM
// walks its items; partial sums leave TMEM through fp32 red.global.add (dW is pre-zeroed by the caller).
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue; two 128-column accumulators so the epilogue of item i overlaps item i+1.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int kBoxBytes = 128 * 64 * 2;     // one [128 pixels x 64 channels] 16-bit box
// A stage = 2 G boxes (co 0..127) + up to 4 A boxes (ci chunk of up to 256): `stage_bytes` / `n_stages` are per plan -- 3 x 64 KB
// when no work item has more than 128 input channels, 2 x 96 KB for N = 256 items.  Why N = 256 where the layer has the channels:
// every loaded tile is used by ONE group of 8 MMAs, so the shared-memory port carries the TMA writes AND the operand reads of the
// same bytes; per 128-pixel patch that is 64 KB written + 64 KB read for 512 tensor cycles at N = 128 (256 B/clk against a
// 128 B/clk port), 96 KB + 96 KB for 1024 tensor cycles at N = 256 (192 B/clk).
constexpr int kMaxStages = 4;

struct WMaps {
    CUtensorMap a[RNR_MAX_VIEWS];
    CUtensorMap g[4];
};

struct WorkItem {          // 8 ints
    int tap, co0, ci_off, n_ci_box, ci_valid, patch_begin, patch_end, n_mma;   // n_mma: N of the MMA (columns of the accumulator)
};

// kPair = 1: CTA pair (tcgen05.mma.cta_group::2, M = 256 output channels; see tc_ptx.cuh).  Both CTAs of a 2-cluster walk the same work
// items (tap, 256 output channels, input-channel chunk, pixel range): CTA r loads the G boxes of ITS 128 output channels and HALF
// of the activation boxes, rank 0 issues the MMAs.  Per 128-pixel patch a CTA then pulls 48 KB instead of 64 KB through its
// operand ring -- the kernel is bound by ring capacity x TMA latency (192 KB cover ~1500 cycles at 128 B/clk, no more), so fewer
// bytes per FLOP is what speeds it up; the stages shrink to 48 KB and a fourth one fits.
template <int kPair>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WMaps maps, const WgradParams p, const WorkItem* __restrict__ work, int n_work,
                int th, int tw, int tiles_y, int tiles_x, int vec, int swap, int stage_bytes, int n_stages, int acc_cols) {
    constexpr int pair = kPair;
    pdl_launch_dependents();
    const int crank = pair ? (int)cluster_ctarank() : 0;
    const int cid = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // work items are walked per cluster
    const int ncl = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* aux = smem + (size_t)n_stages * stage_bytes;
    uint64_t* full_bar = (uint64_t*)aux;
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int v = 0; v < RNR_MAX_VIEWS; v++)
            if (p.aviews[v].ptr) tma_prefetch_desc(&maps.a[v]);
        for (int v = 0; v < 4; v++)
            if (p.gviews[v].ptr) tma_prefetch_desc(&maps.g[v]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], pair ? 8 : 4); }
        fence_barrier_init();
    }
    if (warp == 2) { if (pair) tmem_alloc_cg2(tmem_slot, (uint32_t)(2 * acc_cols)); else tmem_alloc(tmem_slot, (uint32_t)(2 * acc_cols)); }
    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();            // the peer's barriers must exist before rank 0's commits / the peer's TMA reach them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above (barriers, TMEM, tensor-map prefetch) overlapped the predecessor's tail

    if (warp == 0) {
        // TMA producer: the whole warp runs the loop, one elected lane issues (operands stay warp-uniform -> uniform registers)
        int stage = 0;
        uint32_t phase = 0;
        for (int w = cid; w < n_work; w += ncl) {
            const WorkItem wi = work[w];
            const rnr_wtap_t tap = p.taps[wi.tap];
            for (int pt = wi.patch_begin; pt < wi.patch_end; pt++) {
                const int tx_ = pt % tiles_x, ty_ = (pt / tiles_x) % tiles_y, n_ = pt / (tiles_x * tiles_y);
                const int x0 = tx_ * tw, y0 = ty_ * th;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one_sync()) {
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    if (pair) {
                        // my 128 output channels of G + my half of the activation boxes; all bytes of the pair land on rank 0's barrier
                        const int hb = wi.n_ci_box >> 1;
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * (2 + hb) * kBoxBytes));
                        tma_load_4d_cg2(&maps.g[tap.gview], &full_bar[stage], st, wi.co0 + crank * 128, x0, y0, n_);
                        tma_load_4d_cg2(&maps.g[tap.gview], &full_bar[stage], st + kBoxBytes, wi.co0 + crank * 128 + 64, x0, y0, n_);
                        for (int b = 0; b < hb; b++)
                            tma_load_4d_cg2(&maps.a[tap.view], &full_bar[stage], st + (size_t)(2 + b) * kBoxBytes,
                                            tap.c0 + wi.ci_off + 64 * (crank * hb + b), x0 + tap.dx, y0 + tap.dy, n_);
                    } else if (!swap) {
                        mbar_expect_tx(&full_bar[stage], (uint32_t)((2 + wi.n_ci_box) * kBoxBytes));
                        tma_load_4d(&maps.g[tap.gview], &full_bar[stage], st, wi.co0, x0, y0, n_);
                        tma_load_4d(&maps.g[tap.gview], &full_bar[stage], st + kBoxBytes, wi.co0 + 64, x0, y0, n_);
                        for (int b = 0; b < wi.n_ci_box; b++)
                            tma_load_4d(&maps.a[tap.view], &full_bar[stage], st + (size_t)(2 + b) * kBoxBytes,
                                        tap.c0 + wi.ci_off + 64 * b, x0 + tap.dx, y0 + tap.dy, n_);
                    } else {
                        // swapped roles: the 128 input channels of the block are the M operand (boxes 0-1), the output channels the
                        // N operand (boxes 2..): layers with few output channels (64 / 78) then fill all 128 accumulator lanes
                        mbar_expect_tx(&full_bar[stage], (uint32_t)((2 + wi.n_ci_box) * kBoxBytes));
                        tma_load_4d(&maps.a[tap.view], &full_bar[stage], st, tap.c0 + wi.ci_off, x0 + tap.dx, y0 + tap.dy, n_);
                        tma_load_4d(&maps.a[tap.view], &full_bar[stage], st + kBoxBytes, tap.c0 + wi.ci_off + 64, x0 + tap.dx, y0 + tap.dy, n_);
                        for (int b = 0; b < wi.n_ci_box; b++)
                            tma_load_4d(&maps.g[tap.gview], &full_bar[stage], st + (size_t)(2 + b) * kBoxBytes, 64 * b, x0, y0, n_);
                    }
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && !(pair && crank != 0)) {
        // MMA issuer: same structure (see tc_ptx.cuh::elect_one_sync); rank 0 only in a CTA pair
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const uint32_t smem0 = smem_u32(smem);
        for (int w = cid; w < n_work; w += ncl, it++) {
            const WorkItem wi = work[w];
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const uint32_t idesc = swap ? make_idesc(128, wi.n_mma, p.a_dtype, p.g_dtype, 1, 1)
                                        : make_idesc(pair ? 256 : 128, wi.n_mma, p.g_dtype, p.a_dtype, 1, 1);
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
            uint32_t accum = 0;
            for (int pt = wi.patch_begin; pt < wi.patch_end; pt++) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sbase = smem0 + (uint32_t)stage * (uint32_t)stage_bytes;
                if (elect_one_sync()) {
                    const uint64_t dg = make_mnmajor_desc(sbase, kBoxBytes);
                    const uint64_t da = make_mnmajor_desc(sbase + 2 * kBoxBytes, kBoxBytes);
                    if (pair) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            umma_f16_cg2(d_tmem, dg + (uint64_t)(k * (2048 >> 4)), da + (uint64_t)(k * (2048 >> 4)), idesc, accum);
                            accum = 1;
                        }
                        umma_commit_cg2(&empty_bar[stage], 3);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; k++) {     // 128 pixels = 8 MMAs of K=16 (2 KB of rows each)
                            umma_f16(d_tmem, dg + (uint64_t)(k * (2048 >> 4)), da + (uint64_t)(k * (2048 >> 4)), idesc, accum);
                            accum = 1;
                        }
                        umma_commit(&empty_bar[stage]);
                    }
                }
                __syncwarp();
                accum = 1;
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) { if (pair) umma_commit_cg2(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]); }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int w = cid; w < n_work; w += ncl, it++) {
            const WorkItem wi = work[w];
            const rnr_wtap_t tap = p.taps[wi.tap];
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int co = wi.co0 + crank * 128 + row;
            float* dst = p.dw + (int64_t)co * p.s_co + (int64_t)(tap.ci0 + wi.ci_off) * p.s_ci + tap.off;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols);
            if (swap) {
                // lane = input channel (ci_off + row), column = output channel: lanes are adjacent in the ci-contiguous scratch,
                // so every column is one coalesced 128-byte reduction per warp
                const int ci = wi.ci_off + row;
                float* d2 = p.dw + (int64_t)(tap.ci0 + ci) * p.s_ci + tap.off;
                for (int c0 = 0; c0 < wi.n_mma; c0 += 16) {
                    uint32_t rv[16];
                    tmem_ld16(taddr + (uint32_t)c0, rv);
                    tmem_ld_wait();
                    if (ci < tap.nci) {
#pragma unroll
                        for (int e = 0; e < 16; e++)
                            if (c0 + e < p.cout) atomicAdd(d2 + (int64_t)(c0 + e) * p.s_co, __uint_as_float(rv[e]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);        // (swapped roles never run as a pair)
                continue;
            }
            const int ncols = wi.n_ci_box * 64;
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t rv[32];
                tmem_ld16(taddr + (uint32_t)c0, rv);
                if (c0 + 16 < ncols) tmem_ld16(taddr + (uint32_t)(c0 + 16), rv + 16);
                tmem_ld_wait();
                if (co < p.cout) {
                    if (vec) {
                        // dW scratch is ci-contiguous (s_ci == 1): one 128-bit vector reduction per 4 input channels -- 4x fewer
                        // L2 atomic operations than the scalar form, whose lanes (co rows, s_co apart) never share a sector
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int ci = c0 + 4 * j;
                            if (ci + 3 < wi.ci_valid)
                                atomicAdd((float4*)(dst + ci), make_float4(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]),
                                                                           __uint_as_float(rv[4 * j + 2]), __uint_as_float(rv[4 * j + 3])));
                            else
#pragma unroll
                                for (int e = 0; e < 4; e++)
                                    if (ci + e < wi.ci_valid) atomicAdd(dst + ci + e, __uint_as_float(rv[4 * j + e]));
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; e++)
                            if (c0 + e < wi.ci_valid) atomicAdd(dst + (int64_t)(c0 + e) * p.s_ci, __uint_as_float(rv[e]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (pair) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();            // no CTA may exit while its peer can still arrive on its barriers / read its operands
    if (warp == 2) {
        tc_fence_after();
        if (pair) tmem_dealloc_cg2(tmem_base, (uint32_t)(2 * acc_cols)); else tmem_dealloc(tmem_base, (uint32_t)(2 * acc_cols));
    }
}

}  // namespace

int rnr_wgrad_tc_prepare(rnr_wgrad_plan* pl, const rnr_wgrad_problem_t* prob) {
    RNR_REQUIRE((prob->a_dtype == RNR_F16 || prob->a_dtype == RNR_BF16) && (prob->g_dtype == RNR_F16 || prob->g_dtype == RNR_BF16),
                "wgrad_tc: operands must be 16-bit");
    // pixel patches
    int tw = 16;
    while (tw > 1 && tw / 2 >= prob->mX) tw /= 2;
    const int th = 128 / tw;
    pl->tw = tw; pl->th = th;
    pl->tiles_y = rnr_cdiv(prob->mY, th);
    pl->tiles_x = rnr_cdiv(prob->mX, tw);
    const int n_patches = prob->mN * pl->tiles_y * pl->tiles_x;
    for (int i = 0; i < prob->n_aviews; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_a[i], prob->aviews[i], prob->a_dtype, 64, tw, th);
        if (rc) return rc;
    }
    for (int i = 0; i < prob->n_gviews; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_g[i], prob->gviews[i], prob->g_dtype, 64, tw, th);
        if (rc) return rc;
    }
    // output tiles
    // operand roles: M = output channels (blocks of 128), N = input channels (chunks of <= 128) -- or swapped when that wastes
    // fewer accumulator lanes (cout <= 128 only: 64- and 78-channel layers at full resolution)
    long long cost_n = 0, cost_s = 0;
    for (int t = 0; t < prob->n_taps; t++) {
        const int nci = prob->taps[t].nci;
        cost_n += (long long)rnr_cdiv(prob->cout, 128) * 128 * rnr_cdiv(nci, 64) * 64;
        cost_s += (long long)rnr_cdiv(nci, 128) * 128 * rnr_cdiv(prob->cout, 16) * 16;
    }
    int swap = (prob->cout <= 128 && cost_s < cost_n) ? 1 : 0;
    { const char* e = getenv("RNR_WGRAD_SWAP"); if (e) swap = (atoi(e) != 0 && prob->cout <= 128) ? 1 : 0; }
    pl->swap = swap;
    struct OT { int tap, co0, ci_off, nbox, valid, n_mma; };
    // OFF by default: measured on B200 (profiles/r02_perf_unet_c32_n256_{0,1}.txt) the N = 256 items lose -- weight gradient 0.748 ->
    // 0.782 ms per view, 512->512 @64^2: 28 -> 34 us -- the two 96 KB stages cover less TMA latency than three 64 KB stages and
    // the items get coarser; RNR_WGRAD_N256=1 enables them (parity-tested).
    const bool wide = !swap && getenv("RNR_WGRAD_N256") && getenv("RNR_WGRAD_N256")[0] == '1';
    // CTA pair: 256 output channels per work item, every input-channel chunk an even number of 64-channel boxes
    // OFF by default: measured on B200 (profiles/r02_perf_unet_c33_wpair_{0,1}.txt) 0.724 vs 0.733 ms per view -- 1024->256 @128^2 gains
    // (42 -> 38 us), the <= 32^2 layers lose (17 -> 20 us: half as many, coarser items); the deep layers are bound by fixed costs
    // and by the fp32 reductions of their partial sums, not by operand bytes.  RNR_WGRAD_PAIR=1 enables it (parity-tested).
    bool pair = !swap && !wide && prob->cout % 256 == 0 && getenv("RNR_WGRAD_PAIR") && getenv("RNR_WGRAD_PAIR")[0] == '1';
    for (int t = 0; t < prob->n_taps && pair; t++)
        if (prob->taps[t].nci % 128 != 0) pair = false;
    pl->tc_pair = pair ? 1 : 0;
    std::vector<OT> tiles;
    for (int t = 0; t < prob->n_taps; t++) {
        const rnr_wtap_t& tp = prob->taps[t];
        if (swap) {
            for (int ci = 0; ci < tp.nci; ci += 128) {
                OT o = {t, 0, ci, rnr_cdiv(prob->cout, 64), prob->cout, rnr_cdiv(prob->cout, 16) * 16};
                tiles.push_back(o);
            }
            continue;
        }
        // input-channel chunk: 256 where the tap has them (N = 256 MMAs, see the note at kMaxStages), else 128
        const int chunk = (wide && tp.nci >= 256) ? 256 : 128;
        for (int co0 = 0; co0 < prob->cout; co0 += (pair ? 256 : 128))
            for (int ci = 0; ci < tp.nci; ci += chunk) {
                const int rem = tp.nci - ci;
                const int nbox = rem >= chunk ? chunk / 64 : rnr_cdiv(rem, 64);
                OT o = {t, co0, ci, nbox, rem > chunk ? chunk : rem, nbox * 64};
                tiles.push_back(o);
            }
    }
    const int slots = pair ? 74 : 148;        // persistent CTAs (pairs) that walk the work list
    // work items per persistent CTA the pixel split aims at.  Every split adds one more fp32 reduction of the whole weight gradient
    // (red.add of splits x the weight bytes), so layers that already have ~one output tile per SM are not split further: measured
    // per layer (profiles/r02_perf_unet_c34_waves{1,2,3}.txt) 512->512 @32^2 17 -> 14 us, 512->512 @64^2 27 -> 24 us at one wave, while
    // the layers with few output tiles (128->128 @256^2: 9 tiles) need the second wave (26 vs 30 us).  RNR_WGRAD_WAVES overrides.
    int waves = tiles.size() >= 100 ? 1 : 2;
    { const char* e = getenv("RNR_WGRAD_WAVES"); if (e && atoi(e) >= 1 && atoi(e) <= 8) waves = atoi(e); }
    int splits = (int)((slots * waves + (int)tiles.size() - 1) / (int)tiles.size());
    if (splits > n_patches) splits = n_patches;
    if (splits < 1) splits = 1;
    const int per = rnr_cdiv(n_patches, splits);
    std::vector<WorkItem> work;
    for (int s = 0; s < splits; s++) {
        const int pb = s * per, pe = (pb + per < n_patches) ? pb + per : n_patches;
        if (pb >= pe) continue;
        for (const OT& o : tiles) {
            WorkItem w = {o.tap, o.co0, o.ci_off, o.nbox, o.valid, pb, pe, o.n_mma};
            work.push_back(w);
        }
    }
    // vector reductions need a ci-contiguous, 16-byte aligned destination for every (tap, ci chunk)
    pl->vec = (prob->s_ci == 1 && prob->s_co % 4 == 0 && ((uintptr_t)prob->dw & 15) == 0) ? 1 : 0;
    for (int t = 0; t < prob->n_taps && pl->vec; t++)
        if (prob->taps[t].off % 4 != 0 || prob->taps[t].ci0 % 4 != 0) pl->vec = 0;
    pl->n_work = (int)work.size();
    RNR_CHECK(cudaMalloc(&pl->d_work_tab, work.size() * sizeof(WorkItem)));
    RNR_CHECK(cudaMemcpy(pl->d_work_tab, work.data(), work.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    int max_nbox = 2;
    for (const OT& o : tiles) max_nbox = o.nbox > max_nbox ? o.nbox : max_nbox;
    pl->tc_stage_bytes = (2 + (pair ? max_nbox / 2 : max_nbox)) * kBoxBytes;
    pl->tc_stages = (200 * 1024) / pl->tc_stage_bytes > kMaxStages ? kMaxStages : (200 * 1024) / pl->tc_stage_bytes;
    pl->tc_acc_cols = max_nbox > 2 ? 256 : 128;
    pl->smem_bytes = pl->tc_stages * pl->tc_stage_bytes + 256 + 1024;
    pl->grid = pair ? 2 * (pl->n_work < 74 ? pl->n_work : 74) : (pl->n_work < 148 ? pl->n_work : 148);
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RNR_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    return 0;
}

int rnr_wgrad_tc_run(const rnr_wgrad_plan* pl, cudaStream_t stream) {
    WMaps maps;
    memcpy(maps.a, pl->tmap_a, sizeof(maps.a));
    memcpy(maps.g, pl->tmap_g, sizeof(maps.g));
    if (pl->tc_pair) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(pl->grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = pl->smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = rnr_pdl_enabled() ? 2 : 1;
        RNR_CHECK(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<1>, maps, pl->p, (const WorkItem*)pl->d_work_tab, pl->n_work, pl->th, pl->tw,
                                     pl->tiles_y, pl->tiles_x, pl->vec, pl->swap, pl->tc_stage_bytes, pl->tc_stages, pl->tc_acc_cols));
    } else {
        RNR_PDL_LAUNCH(wgrad_tc_kernel<0>, pl->grid, kThreads, pl->smem_bytes, stream, maps, pl->p, (const WorkItem*)pl->d_work_tab, pl->n_work,
                       pl->th, pl->tw, pl->tiles_y, pl->tiles_x, pl->vec, pl->swap, pl->tc_stage_bytes, pl->tc_stages, pl->tc_acc_cols);
    }
    RNR_LAUNCH_CHECK();
    return 0;
}
