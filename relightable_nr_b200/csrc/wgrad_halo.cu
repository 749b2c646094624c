// tcgen05 weight-gradient kernel with SHARED-MEMORY HALO REUSE and one TMEM accumulator per tap.
//
//   dW[co, ci, tap] += sum over pixels  G[pix, co] * A[pix + (dx,dy)_tap, ci]
//
// wgrad_tc.cu loads one [128 pixel x 64 channel] activation box AND the gradient boxes per (tap, 128-pixel patch): a 3x3 layer
// pulls 9 x (G + A) through L2->SM per patch and runs at the L2 delivery limit (the 512^2 layers: 128 B/clk/SM requested, ~57
// delivered, tensor pipe 45 % busy).  Here a work item owns a GROUP of taps that read the same activation view:
//   * per 16x8-pixel tile ONE halo box [(16+ey) x (8+ex) pixels x 64 ch] per 64-channel chunk and the gradient boxes are loaded
//     once; tap t is an MMA whose MN-major B descriptor starts (dy*pitch + dx) * 128 B into the halo tile, 8-row K groups
//     (= one image row of 8 pixels) pitch*128 B apart (SWIZZLE_128B is a function of the absolute shared-memory address);
//   * tap t accumulates in its own TMEM columns [t*N, (t+1)*N): up to 4 taps x 128 or 8 taps x 64 input channels = 512 columns;
//   * the groups are: one kernel ROW (3 taps) of a 3x3 convolution, the 2x2 taps of one parity class of a 4x4 stride-2
//     convolution / transposed convolution;
//   * a 128-channel chunk may be made of two 64-channel boxes from DIFFERENT tensors (the two halves of a skip concatenation).
// L2->SM bytes per tile drop from 9 x 64 KB to 3 x 72 KB (3x3) resp. 16 x 64 KB to 4 x 72 KB (4x4).
// Roles as in wgrad_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue (one co row per
// thread, 128-bit vector reductions into the GEMM-order scratch).
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <map>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int TH = 16, TW = 8;
constexpr int kGBox = 128 * 64 * 2;          // one [128 pixels x 64 channels] gradient box
constexpr int kMaxT = 8;

struct WMaps {
    CUtensorMap a[RNR_MAX_VIEWS];
    CUtensorMap g[4];
};

struct HItem {
    int gview, co0, nbox, n_mma, ntaps, tile_begin, tile_end, ox, oy;
    int view[2], c0[2];          // A boxes: (view, first channel inside the view)
    int ci_dst[2];               // dW input-channel index of each box's channel 0
    int nvalid[2];               // valid channels of each box (<= 64)
    int a_off[kMaxT];            // byte offset of tap t inside the halo tile
    int off[kMaxT];              // element offset of tap t in dW
    int pad;
};

__device__ __forceinline__ uint64_t make_mn_halo_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_halo_kernel(const __grid_constant__ WMaps maps, const WgradParams p, const HItem* __restrict__ items, int n_items,
                  int tiles_y, int tiles_x, int pitch, int a_box_stride, int a_box_bytes, int stages, int stage_bytes, int vec,
                  int nbuf) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* aux = smem + (size_t)stages * stage_bytes;
    uint64_t* full_bar = (uint64_t*)aux;            // [4]
    uint64_t* empty_bar = full_bar + 4;             // [4]
    uint64_t* tfull_bar = empty_bar + 4;            // [2]
    uint64_t* tempty_bar = tfull_bar + 2;           // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t buf_cols = nbuf == 2 ? 256u : 0u;

    if (warp == 0 && lane == 0) {
        for (int v = 0; v < RNR_MAX_VIEWS; v++)
            if (p.aviews[v].ptr) tma_prefetch_desc(&maps.a[v]);
        for (int v = 0; v < 4; v++)
            if (p.gviews[v].ptr) tma_prefetch_desc(&maps.g[v]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above (barriers, TMEM, tensor-map prefetch) overlapped the predecessor's tail

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const HItem it = items[w];
            for (int pt = it.tile_begin; pt < it.tile_end; pt++) {
                const int tx_ = pt % tiles_x, ty_ = (pt / tiles_x) % tiles_y, n_ = pt / (tiles_x * tiles_y);
                const int x0 = tx_ * TW, y0 = ty_ * TH;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one_sync()) {
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * kGBox + it.nbox * a_box_bytes));
                    tma_load_4d(&maps.g[it.gview], &full_bar[stage], st, it.co0, x0, y0, n_);
                    tma_load_4d(&maps.g[it.gview], &full_bar[stage], st + kGBox, it.co0 + 64, x0, y0, n_);
                    for (int b = 0; b < it.nbox; b++)
                        tma_load_4d(&maps.a[it.view[b]], &full_bar[stage], st + 2 * kGBox + (size_t)b * a_box_stride, it.c0[b],
                                    x0 + it.ox, y0 + it.oy, n_);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        int stage = 0;
        uint32_t phase = 0;
        int itn = 0;
        const uint32_t smem0 = smem_u32(smem);
        const uint32_t sbo = (uint32_t)pitch * 128u;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, itn++) {
            const HItem it = items[w];
            const int acc = nbuf == 2 ? (itn & 1) : 0;
            const uint32_t acc_phase = nbuf == 2 ? ((itn >> 1) & 1) : (itn & 1);
            const uint32_t idesc = make_idesc(128, it.n_mma, p.g_dtype, p.a_dtype, 1, 1);
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * buf_cols;
            for (int pt = it.tile_begin; pt < it.tile_end; pt++) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sbase = smem0 + (uint32_t)stage * (uint32_t)stage_bytes;
                const uint32_t first = (pt == it.tile_begin) ? 0u : 1u;
                if (elect_one_sync()) {
                    for (int t = 0; t < it.ntaps; t++) {
                        const uint32_t a_t = sbase + 2u * kGBox + (uint32_t)it.a_off[t];
#pragma unroll
                        for (int k = 0; k < 8; k++) {     // 128 pixels = 8 MMAs of K = 16 (two image rows of 8 pixels)
                            const uint64_t dg = make_mnmajor_desc(sbase + (uint32_t)k * 2048u, kGBox);
                            const uint64_t da = make_mn_halo_desc(a_t + (uint32_t)k * 2u * sbo, (uint32_t)a_box_stride, sbo);
                            umma_f16(d_tmem + (uint32_t)(t * it.n_mma), dg, da, idesc, (k == 0) ? first : 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tfull_bar[acc]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: thread = output channel row ----------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int itn = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, itn++) {
            const HItem it = items[w];
            const int acc = nbuf == 2 ? (itn & 1) : 0;
            const uint32_t acc_phase = nbuf == 2 ? ((itn >> 1) & 1) : (itn & 1);
            const int co = it.co0 + row;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * buf_cols;
            for (int t = 0; t < it.ntaps; t++) {
                for (int b = 0; b < it.nbox; b++) {
                    float* dst = p.dw + (int64_t)co * p.s_co + (int64_t)it.ci_dst[b] * p.s_ci + it.off[t];
                    const int nv = it.nvalid[b];
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t rv[32];
                        const uint32_t col = (uint32_t)(t * it.n_mma + b * 64 + c0);
                        tmem_ld16(taddr + col, rv);
                        tmem_ld16(taddr + col + 16, rv + 16);
                        tmem_ld_wait();
                        if (co < p.cout) {
                            if (vec) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int ci = c0 + 4 * j;
                                    if (ci + 3 < nv)
                                        atomicAdd((float4*)(dst + ci), make_float4(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]),
                                                                                   __uint_as_float(rv[4 * j + 2]), __uint_as_float(rv[4 * j + 3])));
                                    else
#pragma unroll
                                        for (int e = 0; e < 4; e++)
                                            if (ci + e < nv) atomicAdd(dst + ci + e, __uint_as_float(rv[4 * j + e]));
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 32; e++)
                                    if (c0 + e < nv) atomicAdd(dst + (int64_t)(c0 + e) * p.s_ci, __uint_as_float(rv[e]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

struct Seg { int view, c0, ci0, nci; };
struct Pos { int gview, dx, dy; long long off; std::vector<Seg> segs; };

}  // namespace

// Sets pl->halo = 1 when the problem is taken by this kernel (0: the caller uses wgrad_tc.cu).
int rnr_wgrad_halo_prepare(rnr_wgrad_plan* pl, const rnr_wgrad_problem_t* prob) {
    pl->halo = 0;
    const char* env = getenv("RNR_WGRAD_HALO");
    // 0: off; 1 (default): where it measured faster than wgrad_tc.cu -- the 512^2 layers (L2->SM bound there: 0.075 -> 0.064,
    // 0.067 -> 0.051, 0.066 -> 0.058, 0.066 -> 0.049, 0.147 -> 0.064 ms) and 4-tap groups from 256^2 (0.041 -> 0.034 ms); 3-tap
    // groups at 256^2 lose (0.028 -> 0.034 ms: 3 items x 49 pixel splits leave one accumulator set per SM and no epilogue
    // overlap) and smaller layers are reduction-bound either way; 2: every layer
    const int mode = env ? atoi(env) : 1;
    if (mode == 0) return 0;
    const long long npix = (long long)prob->mN * prob->mY * prob->mX;
    if (mode == 1 && npix < 256 * 256) return 0;
    if (prob->cout < 1) return 0;
    // ---- tap positions: entries that differ only in the activation segment they read ----
    std::vector<Pos> pos;
    for (int t = 0; t < prob->n_taps; t++) {
        const rnr_wtap_t& tp = prob->taps[t];
        Pos* hit = nullptr;
        for (Pos& q : pos)
            if (q.gview == tp.gview && q.dx == tp.dx && q.dy == tp.dy && q.off == tp.off) { hit = &q; break; }
        if (!hit) { pos.push_back({tp.gview, tp.dx, tp.dy, tp.off, {}}); hit = &pos.back(); }
        hit->segs.push_back({tp.view, tp.c0, tp.ci0, tp.nci});
    }
    for (Pos& q : pos) std::sort(q.segs.begin(), q.segs.end(), [](const Seg& a, const Seg& b) { return a.ci0 < b.ci0; });
    // boxes of <= 64 channels, identical for every tap position that reads the same views
    struct Box { int view, c0, ci_dst, nvalid; };
    auto boxes_of = [](const Pos& q) {
        std::vector<Box> bx;
        for (const Seg& s : q.segs)
            for (int j = 0; j < s.nci; j += 64) bx.push_back({s.view, s.c0 + j, s.ci0 + j, std::min(64, s.nci - j)});
        return bx;
    };
    // ---- groups: positions with the same (gview, views); a 3x3 kernel (9 positions) is split into its rows ----
    struct Group { std::vector<int> members; };
    std::vector<Group> groups;
    {
        std::map<std::vector<int>, std::vector<int>> buckets;            // key: gview + the views/channel offsets of the segments
        for (size_t i = 0; i < pos.size(); i++) {
            std::vector<int> key = {pos[i].gview};
            for (const Seg& s : pos[i].segs) { key.push_back(s.view); key.push_back(s.c0); key.push_back(s.ci0); key.push_back(s.nci); }
            buckets[key].push_back((int)i);
        }
        for (auto& kv : buckets) {
            const std::vector<Box> bx = boxes_of(pos[kv.second[0]]);
            const int n_item = bx.size() >= 2 ? 128 : 64;
            const int tmax = std::min(kMaxT, 512 / n_item);
            if ((int)kv.second.size() <= tmax) { groups.push_back({kv.second}); continue; }
            std::map<int, std::vector<int>> by_dy;
            for (int i : kv.second) by_dy[pos[i].dy].push_back(i);
            for (auto& r : by_dy) {
                if ((int)r.second.size() > tmax) return 0;
                groups.push_back({r.second});
            }
        }
    }
    if (mode == 1 && npix < 512 * 512) {
        size_t tmax_used = 0;
        for (const Group& g : groups) tmax_used = std::max(tmax_used, g.members.size());
        if (tmax_used < 4) return 0;
    }
    int ex = 0, ey = 0;
    for (const Group& g : groups) {
        int dx0 = 1 << 20, dx1 = -(1 << 20), dy0 = 1 << 20, dy1 = -(1 << 20);
        for (int i : g.members) {
            dx0 = std::min(dx0, pos[i].dx); dx1 = std::max(dx1, pos[i].dx);
            dy0 = std::min(dy0, pos[i].dy); dy1 = std::max(dy1, pos[i].dy);
        }
        ex = std::max(ex, dx1 - dx0); ey = std::max(ey, dy1 - dy0);
    }
    if (ex > 4 || ey > 4) return 0;
    const int pitch = TW + ex, rows = TH + ey;
    const int a_box_bytes = rows * pitch * 128;
    const int a_box_stride = (a_box_bytes + 1023) / 1024 * 1024;
    const int stage_bytes = 2 * kGBox + 2 * a_box_stride;
    int stages = (220 * 1024) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) return 0;
    pl->tw = TW; pl->th = TH;
    pl->tiles_y = rnr_cdiv(prob->mY, TH);
    pl->tiles_x = rnr_cdiv(prob->mX, TW);
    const int n_tiles = prob->mN * pl->tiles_y * pl->tiles_x;
    for (int i = 0; i < prob->n_aviews; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_a[i], prob->aviews[i], prob->a_dtype, 64, pitch, rows);
        if (rc) return rc;
    }
    for (int i = 0; i < prob->n_gviews; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_g[i], prob->gviews[i], prob->g_dtype, 64, TW, TH);
        if (rc) return rc;
    }
    // ---- items ----
    std::vector<HItem> base;
    int max_cols = 0;
    for (const Group& g : groups) {
        int dx0 = 1 << 20, dy0 = 1 << 20;
        for (int i : g.members) { dx0 = std::min(dx0, pos[i].dx); dy0 = std::min(dy0, pos[i].dy); }
        const std::vector<Box> bx = boxes_of(pos[g.members[0]]);
        for (int co0 = 0; co0 < prob->cout; co0 += 128)
            for (size_t b0 = 0; b0 < bx.size(); b0 += 2) {
                HItem it;
                memset(&it, 0, sizeof(it));
                it.gview = pos[g.members[0]].gview; it.co0 = co0;
                it.nbox = (int)std::min<size_t>(2, bx.size() - b0);
                it.n_mma = it.nbox * 64;
                it.ntaps = (int)g.members.size();
                it.ox = dx0; it.oy = dy0;
                for (int b = 0; b < it.nbox; b++) {
                    it.view[b] = bx[b0 + b].view; it.c0[b] = bx[b0 + b].c0; it.ci_dst[b] = bx[b0 + b].ci_dst; it.nvalid[b] = bx[b0 + b].nvalid;
                }
                for (int t = 0; t < it.ntaps; t++) {
                    const Pos& q = pos[g.members[t]];
                    it.a_off[t] = ((q.dy - dy0) * pitch + (q.dx - dx0)) * 128;
                    if (q.off > 0x7fffffffLL) return 0;
                    it.off[t] = (int)q.off;
                }
                max_cols = std::max(max_cols, it.ntaps * it.n_mma);
                base.push_back(it);
            }
    }
    if (base.empty() || max_cols > 512) return 0;
    const int nbuf = max_cols <= 256 ? 2 : 1;
    // one wave: at most 148 work units (a 149th unit would double the kernel's makespan: 3 groups x 50 splits measured 2x slower
    // than 3 x 49), pixel ranges as equal as possible
    int splits = 148 / (int)base.size();
    if (splits > n_tiles) splits = n_tiles;
    if (splits < 1) splits = 1;
    const int per = rnr_cdiv(n_tiles, splits);
    std::vector<HItem> work;
    for (int s = 0; s < splits; s++) {
        const int pb = (int)((long long)n_tiles * s / splits), pe = (int)((long long)n_tiles * (s + 1) / splits);
        if (pb >= pe) continue;
        for (HItem it : base) { it.tile_begin = pb; it.tile_end = pe; work.push_back(it); }
    }
    (void)per;
    pl->vec = (prob->s_ci == 1 && prob->s_co % 4 == 0 && ((uintptr_t)prob->dw & 15) == 0) ? 1 : 0;
    for (const HItem& it : work) {
        for (int t = 0; t < it.ntaps && pl->vec; t++)
            if (it.off[t] % 4 != 0) pl->vec = 0;
        for (int b = 0; b < it.nbox && pl->vec; b++)
            if (it.ci_dst[b] % 4 != 0) pl->vec = 0;
    }
    pl->n_work = (int)work.size();
    RNR_CHECK(cudaMalloc(&pl->d_work_tab, work.size() * sizeof(HItem)));
    RNR_CHECK(cudaMemcpy(pl->d_work_tab, work.data(), work.size() * sizeof(HItem), cudaMemcpyHostToDevice));
    pl->stages = stages;
    pl->halo_pitch = pitch; pl->halo_a_stride = a_box_stride; pl->halo_a_bytes = a_box_bytes; pl->halo_stage_bytes = stage_bytes;
    pl->halo_nbuf = nbuf;
    pl->smem_bytes = stages * stage_bytes + 256 + 1024;
    pl->grid = pl->n_work < 148 ? pl->n_work : 148;
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    pl->halo = 1;
    return 0;
}

int rnr_wgrad_halo_run(const rnr_wgrad_plan* pl, cudaStream_t stream) {
    WMaps maps;
    memcpy(maps.a, pl->tmap_a, sizeof(maps.a));
    memcpy(maps.g, pl->tmap_g, sizeof(maps.g));
    RNR_PDL_LAUNCH(wgrad_halo_kernel, pl->grid, kThreads, pl->smem_bytes, stream, maps, pl->p, (const HItem*)pl->d_work_tab, pl->n_work, pl->tiles_y,
                                                                     pl->tiles_x, pl->halo_pitch, pl->halo_a_stride, pl->halo_a_bytes,
                                                                     pl->stages, pl->halo_stage_bytes, pl->vec, pl->halo_nbuf);
    RNR_LAUNCH_CHECK();
    return 0;
}
