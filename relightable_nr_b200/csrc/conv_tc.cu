// tcgen05 / TMEM / TMA implementation of the generic implicit-GEMM convolution problem.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer   (cp.async.bulk.tensor 4-D box per K-step for A, 2-D box for Wmat)
//   warp 1   : MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=bn, K=16, fp32 accum in TMEM)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue       (tcgen05.ld -> bias/tanh -> global store, per-tile BatchNorm partial sums)
// smem ring of `stages` {A,B} tiles guarded by full/empty mbarriers; the 512 TMEM columns hold two
// 128 x 256 fp32 accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// A K-step is `bk` consecutive channels of one (view, tap): the implicit-GEMM gather is entirely
// expressed by the host-built k-step table, so the same kernel serves Conv2d 3x3/s1, 4x4/s2
// (parity views), ConvTranspose2d 4x4/s2 (4 parity sub-problems) and all of their data gradients.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>

namespace {

constexpr int kThreads = 256;
constexpr int kMaxKsteps = 512;

struct TcMaps {
    CUtensorMap a[RNR_MAX_VIEWS];
    CUtensorMap b;
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const ConvParams p, int bn, int tiles_n, int stages,
               int a_stage_bytes, int b_stage_bytes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)stages * a_stage_bytes;
    uint8_t* aux = smem_b + (size_t)stages * b_stage_bytes;
    uint64_t* full_bar = (uint64_t*)aux;                 // [stages]
    uint64_t* empty_bar = full_bar + stages;             // [stages]
    uint64_t* tfull_bar = empty_bar + stages;            // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    rnr_kstep_t* s_ksteps = (rnr_kstep_t*)(tmem_slot + 4);          // [n_ksteps]
    float* s_part = (float*)(s_ksteps + kMaxKsteps);                // [2 acc][4 warps][2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * tiles_n;

    for (int i = threadIdx.x; i < p.n_ksteps; i += kThreads) s_ksteps[i] = p.ksteps[i];

    if (warp == 0 && lane == 0) {
        for (int v = 0; v < RNR_MAX_VIEWS; v++)
            if (p.views[v].ptr) tma_prefetch_desc(&maps.a[v]);
        tma_prefetch_desc(&maps.b);
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

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int tile_m = t / tiles_n, tile_n = t % tiles_n;
                const int tx_ = tile_m % p.tiles_x, ty_ = (tile_m / p.tiles_x) % p.tiles_y, n_ = tile_m / (p.tiles_x * p.tiles_y);
                const int x0 = tx_ * p.tw, y0 = ty_ * p.th, n0 = tile_n * bn;
                for (int j = 0; j < p.n_ksteps; j++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const rnr_kstep_t ks = s_ksteps[j];
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(128 * p.bk * 2 + bn * p.bk * 2));
                    tma_load_4d(&maps.a[ks.view], &full_bar[stage], smem_a + (size_t)stage * a_stage_bytes, ks.c0, x0 + ks.dx, y0 + ks.dy, n_);
                    tma_load_2d(&maps.b, &full_bar[stage], smem_b + (size_t)stage * b_stage_bytes, j * p.bk, n0);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, bn, p.ab_dtype, p.ab_dtype, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int j = 0; j < p.n_ksteps; j++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_kmajor_desc(smem_u32(smem_a + (size_t)stage * a_stage_bytes), p.bk);
                    const uint64_t db = make_kmajor_desc(smem_u32(smem_b + (size_t)stage * b_stage_bytes), p.bk);
                    const int nk = p.bk >> 4;
                    for (int k = 0; k < nk; k++)
                        umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((j | k) != 0));
                    umma_commit(&empty_bar[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int ry = r / p.tw, rx = r % p.tw;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int tile_m = t / tiles_n, tile_n = t % tiles_n;
            const int tx_ = tile_m % p.tiles_x, ty_ = (tile_m / p.tiles_x) % p.tiles_y, n_ = tile_m / (p.tiles_x * p.tiles_y);
            const int y = ty_ * p.th + ry, x = tx_ * p.tw + rx, n0 = tile_n * bn;
            const bool valid = (y < p.mY && x < p.mX);
            const int64_t obase = (int64_t)n_ * p.out_sn + (int64_t)(y * p.out_my + p.out_py) * p.out_sy +
                                  (int64_t)(x * p.out_mx + p.out_px) * p.out_sx;
            float* part = s_part + (size_t)(acc * 4 + q) * 512;

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            for (int c0 = 0; c0 < bn; c0 += 16) {
                uint32_t rv[16];
                tmem_ld16(taddr + (uint32_t)c0, rv);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    float f = __uint_as_float(rv[e]);
                    const int co = n0 + c0 + e;
                    if (p.epi & RNR_EPI_BIAS) f += (co < p.cout) ? p.bias[co] : 0.f;
                    if (p.epi & RNR_EPI_TANH) f = tanhf(f);
                    v[e] = f;
                }
                if (valid) {
                    if (n0 + c0 + 16 <= p.cout) {
                        if (p.out_dtype == RNR_F32) {
                            float4* o = (float4*)((float*)p.out + obase + n0 + c0);
#pragma unroll
                            for (int e = 0; e < 4; e++) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        } else {
                            __align__(16) unsigned short h[16];
#pragma unroll
                            for (int e = 0; e < 16; e++) h[e] = f2b16(v[e], p.out_dtype);
                            uint4* o = (uint4*)((unsigned short*)p.out + obase + n0 + c0);
                            o[0] = ((const uint4*)h)[0];
                            o[1] = ((const uint4*)h)[1];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e++)
                            if (n0 + c0 + e < p.cout) st_out(p.out, obase + n0 + c0 + e, v[e], p.out_dtype);
                    }
                }
                if (p.epi & RNR_EPI_STATS) {
                    float s1[16], s2[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) { s1[e] = valid ? v[e] : 0.f; s2[e] = s1[e] * s1[e]; }
                    const float a = colsum16(s1, lane);
                    const float b = colsum16(s2, lane);
                    if ((lane & 1) == 0) {
                        const int cc = c0 + col16_of_lane(lane);
                        part[cc] = a;
                        part[256 + cc] = b;
                    }
                }
            }
            // TMEM accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);

            if (p.epi & RNR_EPI_STATS) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int e = threadIdx.x - 128;
                const float* pa = s_part + (size_t)(acc * 4) * 512;
                for (int cc = e; cc < bn; cc += 128) {
                    const int co = n0 + cc;
                    if (co < p.cout) {
                        const float a = pa[cc] + pa[512 + cc] + pa[1024 + cc] + pa[1536 + cc];
                        const float b = pa[256 + cc] + pa[768 + cc] + pa[1280 + cc] + pa[1792 + cc];
                        p.stats[((int64_t)tile_m * 2 + 0) * p.ldstats + co] = a;
                        p.stats[((int64_t)tile_m * 2 + 1) * p.ldstats + co] = b;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

int rnr_encode_view_map(CUtensorMap* map, const rnr_view_t& v, int dtype, int box_c, int box_x, int box_y) {
    auto enc = get_encode_fn();
    RNR_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
    const int es = 2;
    cuuint64_t gdim[4] = {(cuuint64_t)v.dim[0], (cuuint64_t)v.dim[1], (cuuint64_t)v.dim[2], (cuuint64_t)v.dim[3]};
    cuuint64_t gstr[3] = {(cuuint64_t)v.stride[1] * es, (cuuint64_t)v.stride[2] * es, (cuuint64_t)v.stride[3] * es};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < 3; i++) {
        if (gstr[i] == 0) gstr[i] = 16;   // degenerate dims (extent 1) still need a legal stride
        RNR_REQUIRE(gstr[i] % 16 == 0, "view stride %d (%llu bytes) is not 16-byte aligned", i + 1, (unsigned long long)gstr[i]);
    }
    RNR_REQUIRE(((uintptr_t)v.ptr & 15) == 0, "view base pointer is not 16-byte aligned");
    const CUtensorMapSwizzle sw = box_c * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (box_c * es == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(map, dtype == RNR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                     const_cast<void*>(v.ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RNR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(view) failed with CUresult %d (dims %d %d %d %d)", (int)r,
                v.dim[0], v.dim[1], v.dim[2], v.dim[3]);
    return 0;
}

int rnr_conv_tc_prepare(rnr_conv_plan* pl, const rnr_conv_problem_t* prob) {
    RNR_REQUIRE(prob->n_ksteps <= kMaxKsteps, "conv_tc: too many k-steps (%d)", prob->n_ksteps);
    RNR_REQUIRE(prob->n_rows_w % 16 == 0, "conv_tc: Wmat rows must be a multiple of 16");
    auto enc = get_encode_fn();
    RNR_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
    const ConvParams& p = pl->p;
    // N tile
    int bn = prob->n_rows_w < 256 ? prob->n_rows_w : 256;
    if (prob->n_rows_w > 256 && prob->n_rows_w % 256 != 0) {
        // pick the largest multiple of 16 <= 256 that divides the row count
        for (bn = 256; bn >= 16; bn -= 16)
            if (prob->n_rows_w % bn == 0) break;
    }
    int tiles_n = rnr_cdiv(prob->n_rows_w, bn);
    while (bn > 64 && bn % 32 == 0 && p.tiles_m * tiles_n < 148) { bn /= 2; tiles_n = rnr_cdiv(prob->n_rows_w, bn); }
    pl->bn = bn;
    pl->tiles_n = tiles_n;
    const int a_stage = 128 * prob->bk * 2;
    const int b_stage = ((bn * prob->bk * 2) + 1023) / 1024 * 1024;
    const int aux = 64 * 8 + 64 + kMaxKsteps * (int)sizeof(rnr_kstep_t) + 2 * 4 * 512 * 4;
    int stages = (200 * 1024 - aux) / (a_stage + b_stage);
    if (stages > 8) stages = 8;
    RNR_REQUIRE(stages >= 2, "conv_tc: not enough shared memory for 2 stages");
    pl->stages = stages;
    pl->smem_bytes = stages * (a_stage + b_stage) + aux + 1024;
    const int total = p.tiles_m * tiles_n;
    pl->grid = total < 148 ? total : 148;
    for (int i = 0; i < prob->n_views; i++) {
        int rc = rnr_encode_view_map(&pl->tmap_a[i], prob->views[i], prob->ab_dtype, prob->bk, prob->tw, prob->th);
        if (rc) return rc;
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)p.ldw, (cuuint64_t)prob->n_rows_w};
        cuuint64_t gstr[1] = {(cuuint64_t)p.ldw * 2};
        cuuint32_t box[2] = {(cuuint32_t)prob->bk, (cuuint32_t)bn};
        cuuint32_t estr[2] = {1, 1};
        RNR_REQUIRE(gstr[0] % 16 == 0 && ((uintptr_t)prob->wmat & 15) == 0, "conv_tc: Wmat is not 16-byte aligned");
        const CUtensorMapSwizzle sw = prob->bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (prob->bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = enc(&pl->tmap_b, prob->ab_dtype == RNR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         2, const_cast<void*>(prob->wmat), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RNR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(Wmat) failed with CUresult %d", (int)r);
    }
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    return 0;
}

int rnr_conv_tc_run(const rnr_conv_plan* pl, cudaStream_t stream) {
    TcMaps maps;
    memcpy(maps.a, pl->tmap_a, sizeof(maps.a));
    maps.b = pl->tmap_b;
    const int a_stage = 128 * pl->p.bk * 2;
    const int b_stage = ((pl->bn * pl->p.bk * 2) + 1023) / 1024 * 1024;
    conv_tc_kernel<<<pl->grid, kThreads, pl->smem_bytes, stream>>>(maps, pl->p, pl->bn, pl->tiles_n, pl->stages, a_stage, b_stage);
    RNR_LAUNCH_CHECK();
    return 0;
}
