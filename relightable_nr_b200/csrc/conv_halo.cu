// tcgen05 implicit-GEMM convolution with SHARED-MEMORY HALO REUSE of the activation operand.
//
// conv_tc.cu loads one [128 pixel x 64 channel] A tile per (tap, channel chunk): a 3x3 convolution re-reads every input
// pixel 9 times from L2, and the kernel runs at the L2->SM delivery limit (~6 TB/s on B200) with the tensor pipe 85 % idle
// (profiles/r01_conv_tc_v0_raw.csv).  Here the K loop is re-grouped: for every (view, 64-channel chunk) ONE TMA box load
// brings the whole halo tile [(16+ey) rows x (8+ex) pixels x 64 ch] (18x10 for 3x3, 17x9 for the 2x2 sub-convolutions of the
// stride-2 / transposed layers) into shared memory, and each tap of the group is an MMA whose A descriptor simply starts
// (dy*(8+ex)+dx) pixels further into that tile:
//      start address += (dy*pitch + dx) * 128 B,   stride between 8-row groups (SBO) = pitch * 128 B,  SWIZZLE_128B.
// A 16x8-pixel output tile makes every 8-row core-matrix group one image row, so the tap shift is a pure address offset
// (the 128B-swizzle XOR is a function of the absolute shared-memory address for both the TMA write and the UMMA read).
// L2->SM traffic of the A operand drops 9x -> 1.4x (3x3) and 4x -> 1.2x (2x2).
//
// Warp roles as in conv_tc.cu: warp 0 TMA producer (A-halo ring + B ring), warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue.  BatchNorm statistics are accumulated per CTA across all of its tiles and written once
// (stats rows = gridDim.x), so the finalize kernel reads <= 148 rows instead of one per tile.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <algorithm>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int kThreads = 640;
// Warp roles.  The warp scheduler prefers the HIGHEST warp id of an SM sub-partition, so the two latency-critical single-issuer
// roles sit on top (warps 10, 11) and the 8 ALU-heavy epilogue warps below them (a tcgen05.mma issuer that shares its
// scheduler with two higher-numbered epilogue warps was measured at ~90 cycles per MMA issue instead of <= 64).
constexpr int kEpiWarps = 16;          // warps 0-15: TMEM lane group = warp % 4 (hardware rule), 16-column quarter = warp / 4
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kAllocWarp = 16;         // TMEM allocate / free
constexpr int kProducerWarp = 18;      // TMA
constexpr int kMmaWarp = 19;           // tcgen05.mma issue
constexpr int TH = 16, TW = 8;          // output tile (pixels): M = 128, one 8-row UMMA group per image row
constexpr int kMaxGroups = 128;
constexpr int kMaxSub = 4;           // sub-problems fused into one launch (the 4 output parities of a stride-2 layer)
constexpr int kMaxTaps = 512;
constexpr int kAStages = 2;
constexpr int kMaxStatC = 1024;      // BatchNorm statistics: output channels per layer (DNR at nf0 = 80 has 640)

struct HaloCfg {           // per-problem constants of the tap pattern (kernel parameter -> constant bank -> uniform registers)
    int a_off[kMaxSub][16];   // byte offset of tap k inside the halo tile (identical for every group of a sub-problem)
    int out_py[kMaxSub], out_px[kMaxSub];
    int nsub, sub_rows;       // sub-problem s uses Wmat rows [s*sub_rows, (s+1)*sub_rows) and groups [s*n_groups, ...)
};

struct HaloMaps {
    CUtensorMap a[RNR_MAX_VIEWS];
    CUtensorMap b;       // 3-D: T taps x bn rows x 64 ch in one box (cluster size 1)
    CUtensorMap b2;      // 2-D: (bn / cluster size) rows x 64 ch, multicast slice of one tap
};

__device__ __forceinline__ uint64_t make_halo_desc(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1); __expf / __fdividef keep the relative error ~1e-6 (tanhf costs ~40 instructions)
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
    return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ long long* g_trace = nullptr;     // profiling only: [cta][role][64] clock64 stamps (set by rnr_debug_set_trace)
__device__ __forceinline__ void trace(long long* base, int role, int& idx) {
    if (base && idx < 64) base[role * 64 + idx++] = clock64();
}

// kPair = 1: CTA-pair variant (cta_group::2).  A separate instantiation, because a kernel that CONTAINS cta_group::2 instructions is
// rejected ("cluster misconfiguration") when launched without a 2-CTA cluster.
template <int kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const ConvParams p, const HaloGroup* __restrict__ groups, int n_groups,
                 const HaloTap* __restrict__ taps, int n_taps, int bn, int tiles_n, int b_stages, int a_stage_bytes,
                 int b_stage_bytes, int pitch, int a_bytes, int dbg, int T, int cs, int gtaps, const HaloCfg hc, const BnFin bnf,
                 const GStats gs) {
    constexpr int pair = kPair;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)kAStages * a_stage_bytes;
    uint8_t* aux = smem_b + (size_t)b_stages * b_stage_bytes;
    uint64_t* afull = (uint64_t*)aux;                    // [kAStages]
    uint64_t* aempty = afull + kAStages;                 // [kAStages]
    uint64_t* bfull = aempty + kAStages;                 // [b_stages]
    uint64_t* bempty = bfull + 8;                        // [b_stages]
    uint64_t* tfull_bar = bempty + 8;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                // [2]
    uint64_t* raw_bar = tempty_bar + 2;                  // [2]: raw tile full / empty (fused BatchNorm-backward statistics)
    uint32_t* tmem_slot = (uint32_t*)(raw_bar + 2);
    HaloGroup* s_groups = (HaloGroup*)(tmem_slot + 4);               // [kMaxGroups]
    HaloTap* s_taps = (HaloTap*)(s_groups + kMaxGroups);             // [kMaxTaps]
    float* s_acc = (float*)(s_taps + kMaxTaps);                      // [2][kMaxStatC] per-CTA BatchNorm sums
    float* s_stage = s_acc + 2 * kMaxStatC;                                   // [128][68] epilogue staging slab (16-byte aligned)
    // (dbg & 1024: register-direct epilogue without statistics -- the staging slab is not allocated, the B ring got its bytes)
    int64_t* s_rowoff = (int64_t*)(s_stage + ((dbg & 1024) ? 0 : 128 * 68));   // [128] output offset of each tile row
    float* s_colp = (float*)(s_rowoff + 128);                        // [2][8][64] per-pass column partial sums
    unsigned short* s_raw = (unsigned short*)(s_colp + 1024);        // [128][64] producer's raw conv output under the current pass (RNR_EPI_GSTATS)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // cluster of `cs` CTAs: same N tile, `cs` adjacent M tiles (rank r takes M tile mg*cs + r); the B (weight) stage is loaded once
    // per cluster -- every CTA fetches 1/cs of its rows and multicasts them to all members.
    const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
    const int m_groups = (p.tiles_m + cs - 1) / cs;
    const int per_sub = m_groups * tiles_n;
    const int total_tiles = hc.nsub * per_sub;              // cluster-level work items: (sub-problem, M group, N tile)
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
    long long* tr = (g_trace && blockIdx.x < 4) ? g_trace + (size_t)blockIdx.x * 4 * 64 : nullptr;
    int ti = 0;
    if (threadIdx.x == 0) { int z = 0; trace(tr, 3, z); }
    if (dbg & 256) pdl_launch_dependents();

    for (int i = threadIdx.x; i < n_groups * hc.nsub; i += kThreads) s_groups[i] = groups[i];
    for (int i = threadIdx.x; i < n_taps; i += kThreads) s_taps[i] = taps[i];
    for (int i = threadIdx.x; i < 2 * kMaxStatC; i += kThreads) s_acc[i] = 0.f;

    if (warp == kProducerWarp && lane == 0) {
        for (int v = 0; v < RNR_MAX_VIEWS; v++)
            if (p.views[v].ptr) tma_prefetch_desc(&maps.a[v]);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kAStages; s++) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        // CTA pair (cta_group::2, see tc_ptx.cuh): only rank 0 issues MMAs and commits, so every "empty" barrier gets ONE (multicast)
        // arrival per use, and rank 0's accumulator-free barrier collects the epilogue warps of both CTAs
        for (int s = 0; s < b_stages; s++) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], pair ? 1u : (uint32_t)cs); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], pair ? 2 * kEpiWarps : kEpiWarps); }
        mbar_init(&raw_bar[0], 64);            // full: the 64 threads of the two raw-tile loader warps
        mbar_init(&raw_bar[1], kEpiWarps);     // empty: one arrival per epilogue warp
        fence_barrier_init();
    }
    if (warp == kAllocWarp) { if (pair) tmem_alloc_cg2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // peers' barriers must exist before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg & 256) {
        // programmatic dependent launch (RNR_PDL=1, experimental): this grid may have been scheduled while its predecessor in the
        // stream was still running -- everything above (barrier init, TMEM allocation, table loads of constant plan data) overlaps
        // the predecessor's tail; no activation / weight byte is touched before the predecessor has completed and flushed
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    if (warp == kProducerWarp) {
        // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        for (int t = cid; t < total_tiles; t += ncl) {
            const int sub = t / per_sub, ts = t - sub * per_sub;
            const int tile_m = (ts / tiles_n) * cs + crank, tile_n = ts % tiles_n;
            const int tx_ = tile_m % p.tiles_x, ty_ = (tile_m / p.tiles_x) % p.tiles_y, n_ = tile_m / (p.tiles_x * p.tiles_y);
            const int x0 = tx_ * TW, y0 = ty_ * TH, n0 = tile_n * bn + sub * hc.sub_rows;
            if (lane == 0) trace(tr, 0, ti);
            for (int g = 0; g < n_groups; g++) {
                const HaloGroup G = s_groups[sub * n_groups + g];
                mbar_wait(&aempty[as], aph ^ 1);
                if (elect_one_sync()) {
                    if (pair) {
                        // both CTAs load their own halo tile; the bytes of both land on rank 0's barrier
                        if (crank == 0) mbar_expect_tx(&afull[as], (uint32_t)(2 * a_bytes));
                        tma_load_4d_cg2(&maps.a[G.view], &afull[as], smem_a + (size_t)as * a_stage_bytes, G.c0, x0 + G.ox, y0 + G.oy, n_);
                    } else if (dbg & 16) mbar_arrive(&afull[as]);
                    else {
                        mbar_expect_tx(&afull[as], (uint32_t)a_bytes);
                        tma_load_4d(&maps.a[G.view], &afull[as], smem_a + (size_t)as * a_stage_bytes, G.c0, x0 + G.ox, y0 + G.oy, n_);
                    }
                }
                __syncwarp();
                if (++as == kAStages) { as = 0; aph ^= 1; }
                for (int k = 0; k < gtaps; k += T) {
                    const int kblk = g * gtaps + k;
                    if ((dbg & 128) && lane == 0) trace(tr, 0, ti);
                    mbar_wait(&bempty[bs], bph ^ 1);
                    if ((dbg & 128) && lane == 0) trace(tr, 0, ti);
                    if (elect_one_sync()) {
                        if (pair) {
                            // my half of the weight rows of T taps (the MMA reads the other half from the peer's shared memory)
                            if (crank == 0) mbar_expect_tx(&bfull[bs], (uint32_t)(T * bn * 128));
                            tma_load_3d_cg2(&maps.b, &bfull[bs], smem_b + (size_t)bs * b_stage_bytes, 0, n0 + crank * (bn >> 1), kblk);
                        } else if (dbg & 4) mbar_arrive(&bfull[bs]);
                        else {
                            mbar_expect_tx(&bfull[bs], (uint32_t)(T * bn * 128));
                            if (cs == 1) tma_load_3d(&maps.b, &bfull[bs], smem_b + (size_t)bs * b_stage_bytes, 0, n0, kblk);
                            else {
                                const int rows = bn / cs;             // my slice of every tap tile, multicast to the whole cluster
                                for (int tt = 0; tt < T; tt++)
                                    tma_load_2d_mc(&maps.b2, &bfull[bs], smem_b + (size_t)bs * b_stage_bytes + (size_t)tt * bn * 128 +
                                                   (size_t)crank * rows * 128, (kblk + tt) * 64, n0 + crank * rows, cmask);
                            }
                        }
                    }
                    __syncwarp();
                    if (++bs == b_stages) { bs = 0; bph ^= 1; }
                }
            }
            if (lane == 0) trace(tr, 0, ti);
        }
    } else if (warp == kMmaWarp && !(pair && crank != 0)) {
        // ================= MMA issuer (whole warp runs the loop, one elected lane issues; rank 0 only in a CTA pair) =================
        const uint32_t idesc = make_idesc(pair ? 256 : 128, bn, p.ab_dtype, p.ab_dtype, 0, 0);
        const uint32_t sbo = (uint32_t)pitch * 128u;
        const uint32_t b_tap_bytes = (uint32_t)(pair ? (bn >> 1) : bn) * 128u;
        const uint32_t a_smem0 = smem_u32(smem_a), b_smem0 = smem_u32(smem_b);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        int it = 0;
        for (int t = cid; t < total_tiles; t += ncl, it++) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int sub = t / per_sub;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            if (lane == 0) trace(tr, 1, ti);
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
            uint32_t accum = 0;
            for (int g = 0; g < n_groups; g++) {
                mbar_wait(&afull[as], aph);
                tc_fence_after();
                if (g == 0 && lane == 0) trace(tr, 1, ti);
                const uint32_t a_base = a_smem0 + (uint32_t)as * (uint32_t)a_stage_bytes;
                for (int k = 0; k < gtaps; k += T) {
                    if ((dbg & 128) && lane == 0) trace(tr, 1, ti);
                    mbar_wait(&bfull[bs], bph);
                    tc_fence_after();
                    if ((dbg & 128) && lane == 0) trace(tr, 1, ti);
                    const uint32_t b_base = b_smem0 + (uint32_t)bs * (uint32_t)b_stage_bytes;
                    if (elect_one_sync()) {
                        if (!(dbg & 8)) {
                            for (int tt = 0; tt < T; tt++) {
                                const uint64_t da = make_halo_desc(a_base + (uint32_t)hc.a_off[sub][k + tt], sbo);
                                const uint64_t db = make_kmajor_desc(b_base + (uint32_t)tt * b_tap_bytes, 64);
                                if (pair) {
#pragma unroll
                                    for (int kk = 0; kk < 4; kk++) {
                                        umma_f16_cg2(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, accum);
                                        accum = 1;
                                    }
                                } else {
#pragma unroll
                                    for (int kk = 0; kk < 4; kk++) {
                                        umma_f16((dbg & 64) ? (d_tmem ^ ((uint32_t)(kk & 1) << 8)) : d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, accum);
                                        accum = 1;
                                    }
                                }
                            }
                        }
                        if (pair) umma_commit_cg2(&bempty[bs], cmask);
                        else if (cs == 1) umma_commit(&bempty[bs]); else umma_commit_mc(&bempty[bs], cmask);
                    }
                    __syncwarp();
                    accum = 1;
                    if (++bs == b_stages) { bs = 0; bph ^= 1; }
                }
                if (elect_one_sync()) { if (pair) umma_commit_cg2(&aempty[as], cmask); else umma_commit(&aempty[as]); }
                __syncwarp();
                if (++as == kAStages) { as = 0; aph ^= 1; }
            }
            if (elect_one_sync()) { if (pair) umma_commit_cg2(&tfull_bar[acc], cmask); else umma_commit(&tfull_bar[acc]); }
            __syncwarp();
            if (lane == 0) trace(tr, 1, ti);
        }
    } else if (warp < kEpiWarps) {
        // ================= epilogue (8 warps) =================
        // A lone warp per scheduler issues one dependent instruction every ~4 cycles, so the epilogue is spread over 8 warps:
        // warp w owns TMEM lanes 32*(w%4).. (hardware rule) and column half (w-4)/4 of each 64-column pass.
        // Per pass: tcgen05.ld burst (32 columns per warp, one wait), bias / tanh in registers, the pass parked in shared
        // memory [128 rows][68 floats]; then (a) full-row coalesced global stores (16 lanes per 256-byte row) and (b) per-column
        // BatchNorm sums read column-wise from the staging tile.  The accumulator is handed back to the MMA warp as soon as
        // its last pass is in registers.
        const int q = warp & 3, hcol = warp >> 2;
        const int r = q * 32 + lane;
        const int ry = r / TW, rx = r % TW;
        const int e = threadIdx.x;                             // 0..511
        int it = 0;
        int t3 = 1;
        const bool tr4 = (warp == 0 && lane == 0);
        const bool gstat = gs.nseg > 0;                        // BatchNorm-backward statistics of the producer layer(s) (GStats)
        // plain 16-bit outputs (data gradients): every thread converts its 16 accumulator columns and stores them as 32 contiguous
        // bytes of its pixel row straight from registers -- no staging tile, no epilogue barriers
        // BatchNorm statistics in that mode: warp-shuffle column sums (colsum16) into four private accumulator sets -- one per
        // 32-row group, aliased onto the unused staging tile -- added in a fixed order at the end (deterministic, no barrier).
        const bool direct = (dbg & 512) && !(p.epi & (RNR_EPI_BIAS | RNR_EPI_TANH)) && !gstat && p.out_dtype != RNR_F32;
        const bool direct_stats = direct && (p.epi & RNR_EPI_STATS);
        float* s_accq = s_stage;                               // [4 row groups][2][kMaxStatC]
        if (direct_stats) {
            for (int i = e; i < 4 * 2 * kMaxStatC; i += kEpiThreads) s_accq[i] = 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
        uint32_t rph = 0;
        for (int t = cid; t < total_tiles; t += ncl, it++) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int sub = t / per_sub, ts = t - sub * per_sub;
            const int tile_m = (ts / tiles_n) * cs + crank, tile_n = ts % tiles_n;
            const int tx_ = tile_m % p.tiles_x, ty_ = (tile_m / p.tiles_x) % p.tiles_y, n_ = tile_m / (p.tiles_x * p.tiles_y);
            const int y = ty_ * TH + ry, x = tx_ * TW + rx, n0 = tile_n * bn;
            const bool valid = (y < p.mY && x < p.mX && tile_m < p.tiles_m);
            const int64_t my_rowoff = (int64_t)n_ * p.out_sn + (int64_t)(y * p.out_my + hc.out_py[sub]) * p.out_sy +
                                      (int64_t)(x * p.out_mx + hc.out_px[sub]) * p.out_sx;
            if (hcol == 0 && !direct)      // output element offset of every tile row (-1: outside the image)
                s_rowoff[r] = valid ? ((int64_t)n_ * p.out_sn + (int64_t)(y * p.out_my + hc.out_py[sub]) * p.out_sy +
                                       (int64_t)(x * p.out_mx + hc.out_px[sub]) * p.out_sx) : (int64_t)-1;

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (tr4) trace(tr, 2, ti);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            for (int c0 = 0; c0 < ((dbg & 32) ? 0 : bn); c0 += 64) {
                const int pw = min(64, bn - c0);                 // pass width (multiple of 16)
                const int mc0 = hcol * 16;                       // my first column inside the pass
                const int mw = (mc0 < pw) ? 16 : 0;              // my width: 16 columns (pw is a multiple of 16) or nothing
                uint32_t rv[16];
                if (mw > 0) tmem_ld16(taddr + (uint32_t)(c0 + mc0), rv);
                tmem_ld_wait();
                if (tr4 && it == 0) trace(tr, 3, t3);
                if (c0 + 64 >= bn) {
                    // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (pair) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
                }
                if (direct) {
                    if (direct_stats && mw > 0) {
                        // column sums over this warp's 32 rows: lane pair (l, l^1) ends up with column col16_of_lane(l)
                        float a1[16], a2[16];
#pragma unroll
                        for (int k = 0; k < 16; k++) { const float f = valid ? __uint_as_float(rv[k]) : 0.f; a1[k] = f; a2[k] = f * f; }
                        const float s1 = colsum16(a1, lane), s2 = colsum16(a2, lane);
                        const int co = n0 + c0 + mc0 + col16_of_lane(lane);
                        if (!(lane & 1) && co < p.cout && co < kMaxStatC) {
                            s_accq[(q * 2 + 0) * kMaxStatC + co] += s1;        // (q, hcol) is unique per warp: no other writer
                            s_accq[(q * 2 + 1) * kMaxStatC + co] += s2;
                        }
                    }
                    if (mw > 0 && valid) {
                        const int cbase = n0 + c0 + mc0;
                        unsigned short* op = (unsigned short*)p.out + my_rowoff + cbase;
                        if (cbase + 16 <= p.cout) {
                            __align__(16) unsigned short h[16];
#pragma unroll
                            for (int k = 0; k < 16; k++) h[k] = f2b16(__uint_as_float(rv[k]), p.out_dtype);
                            *(uint4*)op = *(const uint4*)h;
                            *(uint4*)(op + 8) = *(const uint4*)(h + 8);
                        } else {
                            for (int k = 0; k < 16; k++)
                                if (cbase + k < p.cout) op[k] = f2b16(__uint_as_float(rv[k]), p.out_dtype);
                        }
                    }
                    continue;
                }
                float* srow = s_stage + r * 68 + mc0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j * 4 < mw) {
                        float v[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float f = __uint_as_float(rv[j * 4 + k]);
                            const int co = n0 + c0 + mc0 + j * 4 + k;
                            if (p.epi & RNR_EPI_BIAS) f += (co < p.cout) ? p.bias[co] : 0.f;
                            if (p.epi & RNR_EPI_TANH) f = fast_tanh(f);
                            v[k] = valid ? f : 0.f;
                        }
                        *(float4*)(srow + j * 4) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (tr4 && it == 0) trace(tr, 3, t3);
                // ---- (a) coalesced stores: 256 threads sweep the [128 x pw] pass row-major ----
                if (!(dbg & 1)) {
                    const bool f32 = (p.out_dtype == RNR_F32);
                    const int ppr = f32 ? (pw >> 2) : (pw >> 3);     // 16-byte pieces per row: 4 floats or 8 halves
                    const int epp = f32 ? 4 : 8;
                    const int npieces = 128 * ppr;
                    const int sh = (ppr == 16) ? 4 : (ppr == 8) ? 3 : (ppr == 4) ? 2 : (ppr == 2) ? 1 : -1;
                    if (f32 && pw == 64 && n0 + c0 + 64 <= p.cout) {
                        // 512 threads x 4 rows: thread e owns the 16-byte piece (e & 15) of rows (e >> 4) + 32 k
                        const int pc = (e & 15) * 4, r0_ = e >> 4;
                        int64_t rb[4];
                        float4 v[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) rb[k] = s_rowoff[r0_ + 32 * k];
#pragma unroll
                        for (int k = 0; k < 4; k++) v[k] = *(const float4*)(s_stage + (r0_ + 32 * k) * 68 + pc);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (rb[k] >= 0) *(float4*)((float*)p.out + rb[k] + n0 + c0 + pc) = v[k];
                    } else
                    for (int i = e; i < npieces; i += kEpiThreads) {
                        const int rr = sh >= 0 ? (i >> sh) : (i / ppr);
                        const int pc = (i - rr * ppr) * epp;
                        const int64_t rb = s_rowoff[rr];
                        if (rb < 0) continue;
                        const int64_t ob = rb + n0 + c0 + pc;
                        const float4 v0 = *(const float4*)(s_stage + rr * 68 + pc);
                        if (f32) {
                            if (n0 + c0 + pc + 4 <= p.cout) *(float4*)((float*)p.out + ob) = v0;
                            else {
                                const float vv[4] = {v0.x, v0.y, v0.z, v0.w};
                                for (int k = 0; k < 4; k++)
                                    if (n0 + c0 + pc + k < p.cout) ((float*)p.out)[ob + k] = vv[k];
                            }
                        } else {
                            const float4 v1 = *(const float4*)(s_stage + rr * 68 + pc + 4);
                            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                            if (n0 + c0 + pc + 8 <= p.cout) {
                                __align__(16) unsigned short h[8];
#pragma unroll
                                for (int k = 0; k < 8; k++) h[k] = f2b16(vv[k], p.out_dtype);
                                *(uint4*)((unsigned short*)p.out + ob) = *(const uint4*)h;
                            } else {
                                for (int k = 0; k < 8; k++)
                                    if (n0 + c0 + pc + k < p.cout) st_out(p.out, ob + k, vv[k], p.out_dtype);
                            }
                        }
                    }
                }
                if (tr4 && it == 0) trace(tr, 3, t3);
                // ---- (b) BatchNorm partial sums: thread e sums column e%64 over rows [16*(e/64), +16) ----
                // deterministic: the 8 row-slice partials of a column are parked in shared memory and added in a fixed order
                // by one thread per column (no floating-point atomics), so two runs give bit-identical statistics
                const bool do_stats = (p.epi & RNR_EPI_STATS) && !(dbg & 2);
                if (do_stats) {
                    const int col = e & 63, part = e >> 6;
                    float a1 = 0.f, a2 = 0.f;
                    if (col < pw) {
                        const float* sp = s_stage + (part * 16) * 68 + col;
#pragma unroll
                        for (int k = 0; k < 16; k++) { const float v = sp[k * 68]; a1 += v; a2 += v * v; }
                    }
                    s_colp[part * 64 + col] = a1;
                    s_colp[512 + part * 64 + col] = a2;
                } else if (gstat) {
                    // gg = g * drop * lrelu'(raw*scale + shift);  column sums of gg and gg * (raw - mean)   (bn_bwd_reduce_fin_kernel).
                    // The producer's raw tile of this pass was staged in shared memory by the loader warps (below).
                    const int col = e & 63, part = e >> 6;
                    const int co = n0 + c0 + col;
                    const bool s1 = gs.nseg > 1 && co >= gs.seg[1].c_lo;
                    const int c_lo = s1 ? gs.seg[1].c_lo : gs.seg[0].c_lo, c_hi = s1 ? gs.seg[1].c_hi : gs.seg[0].c_hi;
                    const int en = s1 ? gs.seg[1].enabled : gs.seg[0].enabled;
                    const bool g_ok = en && col < pw && co >= c_lo && co < c_hi && tile_m < p.tiles_m;
                    float a1 = 0.f, a2 = 0.f;
                    float g_sc = 0.f, g_sh = 0.f, g_mu = 0.f, g_dr = 1.f;
                    if (g_ok) {
                        const int ch = co - c_lo;
                        const float* dp = s1 ? gs.seg[1].drop : gs.seg[0].drop;
                        g_sc = __ldg((s1 ? gs.seg[1].scale : gs.seg[0].scale) + ch);
                        g_sh = __ldg((s1 ? gs.seg[1].shift : gs.seg[0].shift) + ch);
                        g_mu = __ldg((s1 ? gs.seg[1].mean : gs.seg[0].mean) + ch);
                        if (dp) g_dr = __ldg(dp + (int64_t)n_ * (s1 ? gs.seg[1].C : gs.seg[0].C) + ch);
                    }
                    const float g_slope = s1 ? gs.seg[1].slope : gs.seg[0].slope;
                    mbar_wait(&raw_bar[0], rph);
                    if (g_ok) {
                        const float* sp = s_stage + (part * 16) * 68 + col;
                        const unsigned short* rp = s_raw + (part * 16) * 64 + col;
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const float r = cvt16(rp[k * 64], gs.seg[0].raw_dtype);
                            const float z = r * g_sc + g_sh;
                            const float gg = sp[k * 68] * (z > 0.f ? 1.f : g_slope) * g_dr;
                            a1 += gg;
                            a2 += gg * (r - g_mu);
                        }
                    }
                    rph ^= 1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&raw_bar[1]);          // this warp is done with the raw tile
                    s_colp[part * 64 + col] = a1;
                    s_colp[512 + part * 64 + col] = a2;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");      // staging tile is free again
                if (tr4 && it == 0) trace(tr, 3, t3);
                if ((do_stats || gstat) && e < 128) {
                    // threads 0..63: sum of x, threads 64..127: sum of x^2 (the next pass rewrites s_colp only after its own barrier)
                    const int col = e & 63, which = e >> 6;
                    const int co = n0 + c0 + col;
                    if (col < pw && co < p.cout && co < kMaxStatC) {
                        const float* pp = s_colp + which * 512 + col;
                        float a = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; k++) a += pp[k * 64];
                        s_acc[which * kMaxStatC + co] += a;
                    }
                }
            }
            if (dbg & 32) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (pair) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
            }
            if (tr4) trace(tr, 2, ti);
        }
        if (gstat) {
            // this CTA's share of the producers' BatchNorm-backward sums: one fp64 atomic per (channel, sum)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            for (int co = e; co < p.cout && co < kMaxStatC; co += kEpiThreads) {
                const bool s1 = gs.nseg > 1 && co >= gs.seg[1].c_lo;
                const int c_lo = s1 ? gs.seg[1].c_lo : gs.seg[0].c_lo, c_hi = s1 ? gs.seg[1].c_hi : gs.seg[0].c_hi;
                const int en = s1 ? gs.seg[1].enabled : gs.seg[0].enabled;
                if (!en || co < c_lo || co >= c_hi) continue;
                double* tot = s1 ? gs.seg[1].totals : gs.seg[0].totals;
                const int Cs = s1 ? gs.seg[1].C : gs.seg[0].C;
                const float a1 = s_acc[co], a2 = s_acc[kMaxStatC + co];
                if (a1 != 0.f) atomicAdd(tot + (co - c_lo), (double)a1);
                if (a2 != 0.f) atomicAdd(tot + Cs + (co - c_lo), (double)a2);
            }
        }
        if (direct_stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            for (int co = e; co < p.cout && co < kMaxStatC; co += kEpiThreads) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int g = 0; g < 4; g++) { t1 += s_accq[(g * 2 + 0) * kMaxStatC + co]; t2 += s_accq[(g * 2 + 1) * kMaxStatC + co]; }
                s_acc[co] = t1;
                s_acc[kMaxStatC + co] = t2;
            }
        }
        if ((p.epi & RNR_EPI_STATS) && p.stats_tot) {
            // this CTA's sums straight into the layer's fp64 totals: the consumer (rnr_bn_act_fwd_tot) derives mean / invstd itself
            // and no finalize launch sits between the convolution and its activation pass.  (fp64 accumulation of <= 148 fp32
            // partials is exact for any realistic dynamic range, so the result does not depend on the arrival order.)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            for (int co = e; co < p.cout && co < kMaxStatC; co += kEpiThreads) {
                atomicAdd(p.stats_tot + co, (double)s_acc[co]);
                atomicAdd(p.stats_tot + p.cout + co, (double)s_acc[kMaxStatC + co]);
            }
        } else if (p.epi & RNR_EPI_STATS) {
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            for (int co = e; co < p.cout; co += kEpiThreads) {
                p.stats[((int64_t)blockIdx.x * 2 + 0) * p.ldstats + co] = s_acc[co];
                p.stats[((int64_t)blockIdx.x * 2 + 1) * p.ldstats + co] = s_acc[kMaxStatC + co];
            }
            if (bnf.enabled) {
                // ---- fused BatchNorm finalize: last CTA to arrive reduces every CTA's row (same order as bn_finalize_kernel) ----
                int* s_last = (int*)s_rowoff;
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (e == 0) *s_last = (atomicAdd(bnf.ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (*s_last) {
                    __threadfence();
                    // All 512 threads: G = 512 / Cp row groups x Cp channels (Cp = channels rounded up to a power of two); group g sums
                    // rows g, g + G, ... of every CTA's partial sums in double (independent L2 loads, coalesced over channels), the
                    // G partials of a channel are then added in a fixed order: deterministic whichever CTA comes last.
                    double* ss = (double*)s_stage;            // [G][Cp] sums, then [G][Cp] sums of squares: 2 * 512 * 8 B = 8 KB
                    const int T = (int)gridDim.x;
                    for (int c0 = 0; c0 < p.cout; c0 += 512) {
                        const int nch = min(512, p.cout - c0);
                        int Cp = 32;
                        while (Cp < nch) Cp <<= 1;
                        const int G = 512 / Cp;
                        const int g = e / Cp, c = c0 + (e % Cp);
                        double s1 = 0.0, s2 = 0.0;
                        if (c < p.cout) {
#pragma unroll 4
                            for (int t = g; t < T; t += G) {
                                s1 += (double)__ldcg(p.stats + ((int64_t)t * 2 + 0) * p.ldstats + c);
                                s2 += (double)__ldcg(p.stats + ((int64_t)t * 2 + 1) * p.ldstats + c);
                            }
                        }
                        ss[e] = s1; ss[512 + e] = s2;
                        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                        if (e < Cp && c < p.cout) {
                            for (int i = 1; i < G; i++) { s1 += ss[i * Cp + e]; s2 += ss[512 + i * Cp + e]; }
                            const double m = s1 / bnf.count;
                            double var = s2 / bnf.count - m * m;
                            if (var < 0.0) var = 0.0;
                            const float istd = (float)(1.0 / sqrt(var + (double)bnf.eps));
                            const float gm = bnf.gamma ? bnf.gamma[c] : 1.f, b = bnf.beta ? bnf.beta[c] : 0.f;
                            bnf.mean[c] = (float)m;
                            bnf.invstd[c] = istd;
                            bnf.scale[c] = gm * istd;
                            bnf.shift[c] = b - (float)m * gm * istd;
                            if (bnf.running_mean) {
                                const double unbiased = bnf.count > 1.0 ? var * bnf.count / (bnf.count - 1.0) : var;
                                bnf.running_mean[c] = (1.f - bnf.momentum) * bnf.running_mean[c] + bnf.momentum * (float)m;
                                bnf.running_var[c] = (1.f - bnf.momentum) * bnf.running_var[c] + bnf.momentum * (float)unbiased;
                            }
                        }
                        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                    }
                    if (e == 0) {
                        *bnf.ticket = 0;                       // ready for the next launch of this plan
                        if (bnf.num_batches_tracked && bnf.running_mean) *bnf.num_batches_tracked += 1;
                    }
                }
            }
        }
    }

    else if (gs.nseg > 0 && (warp == kAllocWarp || warp == kAllocWarp + 1)) {
        // ================= raw-tile loader (warps 16-17, 64 threads): fused BatchNorm-backward statistics =================
        // For every epilogue pass, in the epilogue's order: the [128 pixel x 64 channel] tile of the producer layer's raw conv
        // output that lies under the pass (reflect-folded for a padded output grid) -> shared memory with 16-byte cp.async, one
        // pass ahead of the epilogue warps, so that the statistics cost them shared-memory reads only.
        const int lt = threadIdx.x - kEpiThreads;               // 0..63
        uint32_t eph = 0;
        for (int t = cid; t < total_tiles; t += ncl) {
            const int sub = t / per_sub, ts = t - sub * per_sub;
            const int tile_m = (ts / tiles_n) * cs + crank, tile_n = ts % tiles_n;
            const int tx_ = tile_m % p.tiles_x, ty_ = (tile_m / p.tiles_x) % p.tiles_y, n_ = tile_m / (p.tiles_x * p.tiles_y);
            const int n0 = tile_n * bn;
            for (int c0 = 0; c0 < bn; c0 += 64) {
                const int pw = min(64, bn - c0);
                const int co0 = n0 + c0;
                const bool s1 = gs.nseg > 1 && co0 >= gs.seg[1].c_lo;
                const int c_lo = s1 ? gs.seg[1].c_lo : gs.seg[0].c_lo, c_hi = s1 ? gs.seg[1].c_hi : gs.seg[0].c_hi;
                const int en = s1 ? gs.seg[1].enabled : gs.seg[0].enabled;
                const int Cs = s1 ? gs.seg[1].C : gs.seg[0].C;
                const unsigned short* rawp = (const unsigned short*)(s1 ? gs.seg[1].raw : gs.seg[0].raw);
                mbar_wait(&raw_bar[1], eph ^ 1);
                if (en && tile_m < p.tiles_m && co0 >= c_lo) {
                    // thread = (pixel column rx of the tile, 16-byte piece pc of the 64 channels); its 16 copies are the 16 tile rows
                    const int ncol = min(pw, c_hi - co0);      // live columns of the pass (multiple of 8: checked on the host)
                    const int rx = lt >> 3, pc = (lt & 7) * 8;
                    if (pc < ncol) {
                        int xx = min(tx_ * TW + rx, p.mX - 1) * p.out_mx + hc.out_px[sub] - gs.pad;
                        xx = xx < 0 ? -xx : (xx >= gs.W ? 2 * gs.W - 2 - xx : xx);
                        const unsigned short* base = rawp + ((int64_t)n_ * gs.H * gs.W + xx) * Cs + (co0 - c_lo) + pc;
                        const int64_t ystride = (int64_t)gs.W * Cs;
                        const uint32_t dst0 = smem_u32(s_raw + rx * 64 + pc);
                        const int ybase = ty_ * TH;
#pragma unroll
                        for (int i = 0; i < TH; i++) {
                            int yy = min(ybase + i, p.mY - 1) * p.out_my + hc.out_py[sub] - gs.pad;
                            yy = yy < 0 ? -yy : (yy >= gs.H ? 2 * gs.H - 2 - yy : yy);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(i * TW * 64 * 2)),
                                         "l"(base + yy * ystride) : "memory");
                        }
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                }
                mbar_arrive(&raw_bar[0]);
                eph ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // no CTA may exit while a peer can still arrive on its barriers
    if (warp == kAllocWarp) {
        tc_fence_after();
        if (pair) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    return fn;
}

}  // namespace

// Returns 0 and sets pl->halo = 1 when the problem fits the halo scheme, 0 with pl->halo = 0 when it does not
// (the caller then uses the generic per-tap kernel), or an error code.
// `nsub` > 1 fuses sub-problems that differ only in their taps' (dx, dy), weight matrix and output parity -- the four output
// parities of a ConvTranspose2d(4, 2, 1) forward or of a Conv2d(4, stride 2) data gradient -- into ONE launch: the tile index
// gets a sub-problem digit.  Their weight matrices must be stacked row-wise in one buffer (sub s at rows [s*n_rows_w, ...)).
int rnr_conv_halo_prepare(rnr_conv_plan* pl, const rnr_conv_problem_t* probs, int nsub) {
    const rnr_conv_problem_t* prob = probs;
    pl->halo = 0;
    const char* env = getenv("RNR_CONV_HALO");
    if (env && env[0] == '0') return 0;
    if (nsub < 1 || nsub > kMaxSub) return 0;
    if (prob->bk != 64) return 0;
    if (prob->cout > kMaxStatC && (prob->epi & RNR_EPI_STATS)) return 0;
    // ---- group the K-steps: maximal runs of consecutive K-steps on the same (view, channel chunk) ----
    // (the engine emits the K-steps chunk-major, so all taps of a chunk are adjacent in K and in Wmat)
    struct Key { int view, c0; };
    std::vector<HaloGroup> groups;
    std::vector<HaloTap> taps;
    int ex = 0, ey = 0, n_groups = 0, gtaps_all = 0;
    std::vector<std::vector<Key>> keys_s(nsub);
    std::vector<std::vector<std::vector<int>>> members_s(nsub);
    for (int sb = 0; sb < nsub; sb++) {
        const rnr_conv_problem_t* pr = probs + sb;
        std::vector<Key>& keys = keys_s[sb];
        std::vector<std::vector<int>>& members = members_s[sb];
        for (int j = 0; j < pr->n_ksteps; j++) {
            const rnr_kstep_t& ks = pr->ksteps[j];
            if (keys.empty() || keys.back().view != ks.view || keys.back().c0 != ks.c0) { keys.push_back({ks.view, ks.c0}); members.emplace_back(); }
            members.back().push_back(j);
        }
        if (sb == 0) { n_groups = (int)keys.size(); gtaps_all = (int)members[0].size(); }
        if ((int)keys.size() != n_groups) return 0;
        for (size_t g = 0; g < keys.size(); g++) {
            if ((int)members[g].size() != gtaps_all) return 0;      // every (view, chunk) group must carry the same taps
            int dx0 = 1 << 20, dx1 = -(1 << 20), dy0 = 1 << 20, dy1 = -(1 << 20);
            for (int j : members[g]) {
                dx0 = std::min(dx0, (int)pr->ksteps[j].dx); dx1 = std::max(dx1, (int)pr->ksteps[j].dx);
                dy0 = std::min(dy0, (int)pr->ksteps[j].dy); dy1 = std::max(dy1, (int)pr->ksteps[j].dy);
            }
            ex = std::max(ex, dx1 - dx0); ey = std::max(ey, dy1 - dy0);
        }
    }
    if (n_groups * nsub > kMaxGroups || prob->n_ksteps > kMaxTaps) return 0;
    if (ex > 4 || ey > 4) return 0;
    if (gtaps_all > 16) return 0;
    const int pitch = TW + ex, rows = TH + ey;
    memset(pl->halo_a_off, 0, sizeof(pl->halo_a_off));
    for (int sb = 0; sb < nsub; sb++) {
        const rnr_conv_problem_t* pr = probs + sb;
        const std::vector<Key>& keys = keys_s[sb];
        const std::vector<std::vector<int>>& members = members_s[sb];
        const size_t tap0 = taps.size();
        for (size_t g = 0; g < keys.size(); g++) {
            int dx0 = 1 << 20, dy0 = 1 << 20;
            for (int j : members[g]) { dx0 = std::min(dx0, (int)pr->ksteps[j].dx); dy0 = std::min(dy0, (int)pr->ksteps[j].dy); }
            HaloGroup G;
            G.view = (int16_t)keys[g].view; G.c0 = (int16_t)keys[g].c0; G.ox = (int16_t)dx0; G.oy = (int16_t)dy0;
            G.first_tap = (int16_t)taps.size(); G.n_taps = (int16_t)members[g].size();
            for (int j : members[g]) {
                HaloTap t;
                t.a_off = ((pr->ksteps[j].dy - dy0) * pitch + (pr->ksteps[j].dx - dx0)) * 128;
                t.wcol = j * 64;
                taps.push_back(t);
            }
            groups.push_back(G);
        }
        // the halo kernel addresses tap k of every group through one offset table: all groups must share the tap pattern
        for (size_t g = 0; g < keys.size(); g++)
            for (int k = 0; k < gtaps_all; k++)
                if (taps[tap0 + g * gtaps_all + k].a_off != taps[tap0 + k].a_off) return 0;
        for (int k = 0; k < gtaps_all; k++) pl->halo_a_off[sb][k] = taps[tap0 + k].a_off;
        pl->halo_out_py[sb] = pr->out_py; pl->halo_out_px[sb] = pr->out_px;
        // sub-problems must agree on everything the kernel takes from sub-problem 0
        if (sb > 0) {
            const rnr_conv_problem_t* p0 = probs;
            if (pr->n_views != p0->n_views || pr->bk != p0->bk || pr->n_ksteps != p0->n_ksteps || pr->n_rows_w != p0->n_rows_w ||
                pr->cout != p0->cout || pr->mN != p0->mN || pr->mY != p0->mY || pr->mX != p0->mX || pr->out != p0->out ||
                pr->out_dtype != p0->out_dtype || pr->out_sn != p0->out_sn || pr->out_sy != p0->out_sy || pr->out_sx != p0->out_sx ||
                pr->out_my != p0->out_my || pr->out_mx != p0->out_mx || pr->epi != p0->epi || pr->bias != p0->bias ||
                pr->ab_dtype != p0->ab_dtype || pr->ldstats != p0->ldstats)
                return 0;
            for (int v = 0; v < p0->n_views; v++)
                if (memcmp(&pr->views[v], &p0->views[v], sizeof(rnr_view_t)) != 0) return 0;
            const size_t sub_bytes = (size_t)p0->n_rows_w * p0->n_ksteps * p0->bk * 2;
            if ((const char*)pr->wmat != (const char*)p0->wmat + sb * sub_bytes) return 0;
        }
    }
    pl->halo_nsub = nsub;
    pl->halo_gtaps = gtaps_all;
    const bool uniform = true;
    std::vector<std::vector<int>> members(1, std::vector<int>(gtaps_all));
    // ---- tiling ----
    ConvParams& p = pl->p;
    p.th = TH; p.tw = TW;
    p.tiles_y = rnr_cdiv(prob->mY, TH); p.tiles_x = rnr_cdiv(prob->mX, TW);
    p.tiles_m = p.tiles_y * p.tiles_x * prob->mN;
    int bn = prob->n_rows_w < 256 ? prob->n_rows_w : 256;
    if (prob->n_rows_w > 256 && prob->n_rows_w % 256 != 0) {
        for (bn = 256; bn >= 16; bn -= 16)
            if (prob->n_rows_w % bn == 0) break;
    }
    int tiles_n = rnr_cdiv(prob->n_rows_w, bn);
    // a wider N tile reads fewer shared-memory operand bytes per FLOP (the SS-mode MMA is shared-memory-bandwidth bound below N = 256),
    // so N is only narrowed when the launch would otherwise leave most SMs idle
    while (bn > 64 && bn % 32 == 0 && p.tiles_m * tiles_n * nsub < (bn > 128 ? 100 : 50)) { bn /= 2; tiles_n = rnr_cdiv(prob->n_rows_w, bn); }
    // <= 32^2 layers: even N = 64 leaves most SMs without a tile and every CTA with a long, latency-bound K loop over a
    // 512-channel input; N = 32 doubles the CTAs and halves the weight bytes each one has to pull through its B ring
    if (bn == 64 && p.tiles_m * tiles_n * nsub <= 74 && prob->n_rows_w % 32 == 0 && !(getenv("RNR_CONV_BN32") && getenv("RNR_CONV_BN32")[0] == '0')) {
        bn = 32; tiles_n = rnr_cdiv(prob->n_rows_w, bn);
    }
    pl->bn = bn;
    pl->tiles_n = tiles_n;
    const int a_stage = ((rows * pitch * 128) + 1023) / 1024 * 1024;
    // epilogue straight from registers (no staging tile): measured faster for every N tile except N = 64 -- there the whole pixel row
    // is one 128-byte line written by four warps, and the staged, row-coalesced stores win by ~3 us per 512^2 launch
    // (profiles/r02_perf_unet_c28_direct{0,1,2}.txt).  RNR_CONV_DIRECT=0 off, =2 for every layer.  Without BatchNorm statistics
    // (data gradients) the staging slab is not needed at all and its 34 KB go to the weight ring.
    int direct = 0;
    { const char* d = getenv("RNR_CONV_DIRECT"); const int want = d ? atoi(d) : 1; if (want >= 2 || (want == 1 && bn != 64)) direct = 1; }
    if ((prob->epi & (RNR_EPI_BIAS | RNR_EPI_TANH)) || prob->out_dtype == RNR_F32) direct = 0;
    const int nostage = direct && !(prob->epi & (RNR_EPI_STATS | RNR_EPI_GSTATS)) && !(getenv("RNR_CONV_NOSTAGE") && getenv("RNR_CONV_NOSTAGE")[0] == '0');
    const int aux_full = 64 * 8 + 64 + kMaxGroups * (int)sizeof(HaloGroup) + kMaxTaps * (int)sizeof(HaloTap) + 2 * kMaxStatC * 4 + 128 * 68 * 4 + 128 * 8 + 1024 * 4 +
                    ((prob->epi & RNR_EPI_GSTATS) ? 128 * 64 * 2 : 0);     // + the raw tile of the fused BatchNorm-backward statistics
    const int aux = aux_full - (nostage ? 128 * 68 * 4 : 0);
    const int budget = 212 * 1024 - aux - kAStages * a_stage;
    // taps per B stage: one mbarrier hand-shake (~400 cycles of latency in the single-thread producer / issuer loops) must
    // cover enough tensor work, so a stage holds T taps = T*4 MMAs; T divides the taps of a group
    const int gtaps = uniform ? (int)members[0].size() : 1;
    // CTA pair (tcgen05.mma.cta_group::2, M = 256): two CTAs on adjacent M tiles of the same N tile share the weight operand -- each
    // loads and holds HALF of its rows, the pair MMA reads both halves.  Per SM that halves the weight bytes pulled from L2 and
    // the shared-memory bytes the tensor core reads for B (the single-CTA SS MMA saturates the 128 B/clk shared-memory port at
    // N <= 128: profiles/r02_mma_issue_bench.txt).  Measured per layer (profiles/r02_perf_unet_c21_pair*.txt): N >= 128 layers gain
    // 10-35 % (1024->256 @128^2: 831 -> 998 TFLOP/s), the N = 64 layers at 512^2 lose ~10 % (their traffic is the activation
    // operand, which a pair cannot share), so the pair is used from N = 128 up.
    // RNR_CONV_PAIR=0 switches it off, RNR_CONV_PAIR_MINBN overrides the threshold.
    int pair = 0;
    {
        const char* pe = getenv("RNR_CONV_PAIR");
        const char* me = getenv("RNR_CONV_PAIR_MINBN");
        const int want = pe ? atoi(pe) : 1;
        const int min_bn = me ? atoi(me) : 128;
        if (want > 0 && p.tiles_m >= 2 && bn % 16 == 0 && bn >= min_bn) pair = 1;
    }
    pl->halo_pair = pair;
    const int bn_cta = pair ? bn / 2 : bn;                 // weight rows held per CTA
    int T = 1;
    for (int cand = gtaps; cand >= 1; cand--) {
        if (gtaps % cand) continue;
        const int stage = cand * bn_cta * 128;
        if (stage <= 72 * 1024 && budget / stage >= 2) { T = cand; break; }
    }
    { const char* te = getenv("RNR_CONV_T"); if (te && atoi(te) >= 1 && gtaps % atoi(te) == 0 && atoi(te) * bn_cta * 128 <= 72 * 1024) T = atoi(te); }
    const int b_stage = T * bn_cta * 128;                  // multiple of 1024 (bn_cta is a multiple of 8 -> 1024 B)
    int b_stages = budget / b_stage;
    if (b_stages > 8) b_stages = 8;
    if (b_stages < 2) return 0;
    pl->halo_T = T;
    // cluster size: `cs` CTAs working on adjacent M tiles can share the weight stream by TMA multicast.  Measured over the 22
    // layers (tools/perf_unet.py, RNR_CONV_CLUSTER = 1 / 2 / 4): forward 0.794 / 0.821 / 1.076 ms, data gradient 0.850 / 0.884 /
    // 1.203 ms -- at this size multicast does not reduce L2 traffic enough to pay for the cluster launch, the cluster barriers
    // and the lock-step B ring, so the default is no cluster.
    int cs = pair ? 2 : 1;
    if (!pair) {
        const char* ce = getenv("RNR_CONV_CLUSTER");
        const int want = ce ? atoi(ce) : 1;
        for (int c = want; c >= 2; c >>= 1)
            if ((bn / c) % 8 == 0 && bn % c == 0 && p.tiles_m >= 2 * c) { cs = c; break; }
    }
    pl->halo_cs = cs;
    pl->stages = b_stages;
    pl->halo_a_stage = a_stage;
    pl->halo_b_stage = b_stage;
    pl->halo_pitch = pitch;
    pl->halo_a_bytes = rows * pitch * 128;
    pl->smem_bytes = kAStages * a_stage + b_stages * b_stage + aux + 1024;
    {
        const int groups_total = rnr_cdiv(p.tiles_m, cs) * tiles_n * nsub;
        const int max_clusters = 148 / cs;
        pl->grid = (groups_total < max_clusters ? groups_total : max_clusters) * cs;
    }
    // ---- tensor maps ----
    auto enc = encode_fn();
    RNR_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
    for (int i = 0; i < prob->n_views; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_a[i], prob->views[i], prob->ab_dtype, 64, pitch, rows);
        if (rc) return rc;
    }
    {
        // Wmat [n_rows, ldw] viewed as (64 channels, n rows, K blocks): one box = T consecutive K blocks = T [bn x 64] tiles
        cuuint64_t gdim[3] = {64, (cuuint64_t)prob->n_rows_w * nsub, (cuuint64_t)(p.ldw / 64)};
        cuuint64_t gstr[2] = {(cuuint64_t)p.ldw * 2, 128};
        cuuint32_t box[3] = {64u, (cuuint32_t)bn_cta, (cuuint32_t)T};
        cuuint32_t estr[3] = {1, 1, 1};
        RNR_REQUIRE(gstr[0] % 16 == 0 && ((uintptr_t)prob->wmat & 15) == 0, "conv_halo: Wmat is not 16-byte aligned");
        CUresult r = enc(&pl->tmap_b, prob->ab_dtype == RNR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         3, const_cast<void*>(prob->wmat), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RNR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(Wmat 3-D) failed with CUresult %d", (int)r);
    }
    if (cs > 1 && !pair) {
        cuuint64_t gdim[2] = {(cuuint64_t)p.ldw, (cuuint64_t)prob->n_rows_w * nsub};
        cuuint64_t gstr[1] = {(cuuint64_t)p.ldw * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)(bn / cs)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&pl->tmap_b2, prob->ab_dtype == RNR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         2, const_cast<void*>(prob->wmat), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RNR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(Wmat slice) failed with CUresult %d", (int)r);
    } else {
        pl->tmap_b2 = pl->tmap_b;
    }
    // ---- device tables ----
    RNR_CHECK(cudaMalloc(&pl->d_groups, sizeof(HaloGroup) * groups.size()));
    RNR_CHECK(cudaMemcpy(pl->d_groups, groups.data(), sizeof(HaloGroup) * groups.size(), cudaMemcpyHostToDevice));
    RNR_CHECK(cudaMalloc(&pl->d_taps, sizeof(HaloTap) * taps.size()));
    RNR_CHECK(cudaMemcpy(pl->d_taps, taps.data(), sizeof(HaloTap) * taps.size(), cudaMemcpyHostToDevice));
    pl->n_groups = n_groups;
    pl->n_taps = (int)taps.size();
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(conv_halo_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RNR_CHECK(cudaFuncSetAttribute(conv_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    { const char* d = getenv("RNR_CONV_DBG"); pl->dbg = d ? atoi(d) : 0; }
    if (rnr_pdl_enabled()) pl->dbg |= 256;      // programmatic dependent launch (common.cuh)
    if (direct) pl->dbg |= 512;
    if (nostage) pl->dbg |= 1024;
    pl->halo = 1;
    return 0;
}

// Fuse the BatchNorm statistics finalize into this plan's launches (halo kernel with RNR_EPI_STATS only; returns
// cudaErrorNotSupported otherwise and the caller keeps the separate rnr_bn_finalize launch).  May be called again at any time
// (e.g. running statistics on / off); `enabled = 0` restores the plain behaviour.
extern "C" int rnr_conv_plan_set_bn(rnr_conv_plan_t* pl, const float* gamma, const float* beta, double count, float eps, float momentum,
                                    float* mean, float* invstd, float* scale, float* shift, float* running_mean, float* running_var,
                                    long long* num_batches_tracked, int* ticket, int enabled) {
    RNR_REQUIRE(pl, "rnr_conv_plan_set_bn: null plan");
    if (!(pl->impl == 1 && pl->halo && (pl->p.epi & RNR_EPI_STATS))) return (int)cudaErrorNotSupported;
    RNR_REQUIRE(!enabled || (mean && invstd && scale && shift && ticket), "rnr_conv_plan_set_bn: null output pointer");
    BnFin& b = pl->bnf;
    b.gamma = gamma; b.beta = beta; b.mean = mean; b.invstd = invstd; b.scale = scale; b.shift = shift;
    b.running_mean = running_mean; b.running_var = running_var; b.num_batches_tracked = num_batches_tracked; b.ticket = ticket;
    b.count = count; b.eps = eps; b.momentum = momentum; b.enabled = enabled ? 1 : 0;
    return 0;
}

// Fuse the BatchNorm-backward statistics of the producer layer(s) of this DATA-GRADIENT plan's output into its epilogue (see
// GStats in conv_internal.cuh).  cudaErrorNotSupported (no error text) when the plan cannot carry them -- the caller keeps the
// separate rnr_bn_bwd_reduce_fin pass.  May be called again (e.g. new dropout-mask pointers); nseg = 0 switches it off.
extern "C" int rnr_conv_plan_set_gstats(rnr_conv_plan_t* pl, const rnr_gstat_seg_t* segs, int nseg, int H, int W, int pad) {
    RNR_REQUIRE(pl, "rnr_conv_plan_set_gstats: null plan");
    if (nseg == 0) { pl->gst.nseg = 0; return 0; }
    if (!(pl->impl == 1 && pl->halo) || (pl->p.epi & RNR_EPI_STATS) || (pl->halo_cs != 1 && !pl->halo_pair)) return (int)cudaErrorNotSupported;
    if (nseg < 1 || nseg > 2 || pl->p.out_dtype == RNR_F32 || pl->p.cout > kMaxStatC) return (int)cudaErrorNotSupported;
    if (!(pl->p.epi & RNR_EPI_GSTATS)) return (int)cudaErrorNotSupported;      // no shared memory reserved for the raw tile
    if (pad != 0 && pad != 1) return (int)cudaErrorNotSupported;
    // the output grid must be the activation's (reflect-padded) plane
    if (pl->p.mY * pl->p.out_my != H + 2 * pad || pl->p.mX * pl->p.out_mx != W + 2 * pad || H < 2 || W < 2) return (int)cudaErrorNotSupported;
    GStats g;
    memset(&g, 0, sizeof(g));
    g.nseg = nseg; g.H = H; g.W = W; g.pad = pad;
    for (int i = 0; i < nseg; i++) {
        const rnr_gstat_seg_t& s = segs[i];
        GStatSeg& d = g.seg[i];
        d.enabled = s.raw != nullptr;
        // a 64-column epilogue pass must lie inside one segment; 16-bit raw tensors only (two values per register)
        if (i > 0 && (s.c_lo % 64 != 0 || pl->bn % 64 != 0 || s.c_lo != segs[i - 1].c_hi)) return (int)cudaErrorNotSupported;
        if (d.enabled && (s.raw_dtype == RNR_F32 || s.raw_dtype != segs[0].raw_dtype || !s.scale || !s.shift || !s.mean || !s.totals ||
                          s.c_hi - s.c_lo > s.C || s.c_lo % 8 != 0 || (s.c_hi - s.c_lo) % 8 != 0 || s.C % 8 != 0 || (int64_t)pl->p.mN * H * W * s.C >= (1ll << 31)))
            return (int)cudaErrorNotSupported;
        d.raw = s.raw; d.scale = s.scale; d.shift = s.shift; d.mean = s.mean; d.drop = s.drop; d.totals = s.totals;
        d.c_lo = s.c_lo; d.c_hi = s.c_hi; d.C = s.C; d.raw_dtype = s.raw_dtype; d.slope = s.slope;
    }
    if (!g.seg[0].enabled && g.seg[1].enabled) g.seg[0].raw_dtype = g.seg[1].raw_dtype;
    if (!g.seg[0].enabled && !(nseg > 1 && g.seg[1].enabled)) g.nseg = 0;
    pl->gst = g;
    return 0;
}

// BatchNorm batch sums of this plan's output as fp64 totals [2, cout] (zero before the first launch; rnr_bn_act_fwd_tot consumes and
// re-zeroes them) instead of per-CTA rows + rnr_bn_finalize.  totals = NULL restores the rows.  cudaErrorNotSupported (no error
// text) when the plan is not a halo-kernel launch with RNR_EPI_STATS.
extern "C" int rnr_conv_plan_set_stat_totals(rnr_conv_plan_t* pl, double* totals) {
    RNR_REQUIRE(pl, "rnr_conv_plan_set_stat_totals: null plan");
    if (!(pl->impl == 1 && pl->halo && (pl->p.epi & RNR_EPI_STATS)) || pl->p.cout > kMaxStatC || pl->bnf.enabled) return (int)cudaErrorNotSupported;
    pl->p.stats_tot = totals;
    return 0;
}

extern "C" int rnr_debug_set_trace(long long* buf) {
    RNR_CHECK(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
    return 0;
}

int rnr_conv_halo_run(const rnr_conv_plan* pl, cudaStream_t stream) {
    HaloMaps maps;
    memcpy(maps.a, pl->tmap_a, sizeof(maps.a));
    maps.b = pl->tmap_b;
    maps.b2 = pl->tmap_b2;
    HaloCfg hc;
    memcpy(hc.a_off, pl->halo_a_off, sizeof(hc.a_off));
    memcpy(hc.out_py, pl->halo_out_py, sizeof(hc.out_py));
    memcpy(hc.out_px, pl->halo_out_px, sizeof(hc.out_px));
    hc.nsub = pl->halo_nsub;
    hc.sub_rows = pl->p.n_rows_w;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(pl->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl->halo_cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pl->dbg & 256) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    if (pl->halo_pair) {
        RNR_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1>, maps, pl->p, (const HaloGroup*)pl->d_groups, pl->n_groups,
                                 (const HaloTap*)pl->d_taps, pl->n_taps, pl->bn, pl->tiles_n, pl->stages, pl->halo_a_stage,
                                 pl->halo_b_stage, pl->halo_pitch, pl->halo_a_bytes, pl->dbg, pl->halo_T, pl->halo_cs, pl->halo_gtaps, hc, pl->bnf,
                                 pl->gst));
    } else {
        RNR_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<0>, maps, pl->p, (const HaloGroup*)pl->d_groups, pl->n_groups,
                                 (const HaloTap*)pl->d_taps, pl->n_taps, pl->bn, pl->tiles_n, pl->stages, pl->halo_a_stage,
                                 pl->halo_b_stage, pl->halo_pitch, pl->halo_a_bytes, pl->dbg, pl->halo_T, pl->halo_cs, pl->halo_gtaps, hc, pl->bnf,
                                 pl->gst));
    }
    rnr_count_launch();
    return 0;
}
