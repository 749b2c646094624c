// Internal structures shared by the SIMT and tcgen05 implementations of the generic
// implicit-GEMM convolution (rnr_conv_*) and weight-gradient (rnr_wgrad_*) problems.
#pragma once
#include "common.cuh"
#include <cuda.h>

struct ConvParams {
    ViewD views[RNR_MAX_VIEWS];
    const rnr_kstep_t* ksteps;   // device
    int n_ksteps, bk, ab_dtype;
    const void* wmat;
    int ldw, n_rows_w, cout;
    int mN, mY, mX, th, tw, tiles_y, tiles_x, tiles_m;
    void* out;
    int out_dtype;
    int64_t out_sn, out_sy, out_sx;
    int out_my, out_mx, out_py, out_px;
    int epi;
    const float* bias;
    float* stats;
    int ldstats;
    double* stats_tot;   // != nullptr: the per-CTA BatchNorm sums go into [2, cout] fp64 totals (atomics) instead of `stats` rows
};

// BatchNorm statistics finalize fused into the halo kernel: the LAST CTA to publish its partial sums (ticket counter) reduces
// all gridDim.x rows in the fixed order of bn_finalize_kernel (bit-identical results, whichever CTA comes last) and writes
// mean / invstd / scale / shift (+ running statistics, num_batches_tracked) -- one launch less on the dependent chain per layer.
struct BnFin {
    const float* gamma;
    const float* beta;
    float* mean;
    float* invstd;
    float* scale;
    float* shift;
    float* running_mean;       // nullptr: no running-statistics update
    float* running_var;
    long long* num_batches_tracked;   // nullptr: not counted here
    int* ticket;               // zero-initialised; reset by the last CTA
    double count;
    float eps, momentum;
    int enabled;
};

// BatchNorm-backward statistics fused into the epilogue of a DATA-GRADIENT launch of the halo kernel.  The launch produces
// d loss / d act for the (one or two concatenated) inputs of a layer; each input is the activation of a producer layer
//     act = drop * lrelu(raw * scale + shift)
// whose BatchNorm backward needs  sum(gg)  and  sum(gg * (raw - mean))  over all pixels,  gg = g * drop * lrelu'(.).  Both are
// linear in g, so every consumer's data-gradient launch adds its own share (the epilogue holds g in fp32, reads `raw` at the
// same -- reflect-folded -- pixel) and the separate reduction pass over the gradient tensor (bn_bwd_reduce_fin) disappears.
struct GStatSeg {
    const void* raw;          // producer's pre-BatchNorm conv output [N, H, W, C]
    const float* scale;
    const float* shift;
    const float* mean;
    const float* drop;        // [N, C] dropout scale (nullptr: none)
    double* totals;           // [2, C]: sum(gg), sum(gg * (raw - mean)); fp64 atomics, zeroed by the consumer of the totals
    int c_lo, c_hi;           // output columns [c_lo, c_hi) of the launch belong to this producer (channel = column - c_lo)
    int C, raw_dtype;
    float slope;
    int enabled;
};
struct GStats {
    GStatSeg seg[2];
    int nseg;
    int H, W;                 // interior size of the activation
    int pad;                  // 1: the launch's output grid is the reflect-padded plane [H+2, W+2]
};

struct HaloGroup {       // one (view, 64-channel chunk): a single halo box load serves all of its taps
    int16_t view, c0, ox, oy, first_tap, n_taps;
};
struct HaloTap {
    int32_t a_off;       // byte offset of the tap's first pixel inside the halo tile
    int32_t wcol;        // first Wmat column (element index) of this K-step
};

struct rnr_conv_plan {
    ConvParams p;
    int impl;
    rnr_kstep_t* d_ksteps;
    // tcgen05 path
    CUtensorMap tmap_a[RNR_MAX_VIEWS];
    CUtensorMap tmap_b;
    int bn;            // N tile of the tensor-core kernel
    int tiles_n;
    int stages;
    int smem_bytes;
    int grid;
    // halo-reuse kernel (conv_halo.cu)
    int halo;
    int halo_a_stage, halo_b_stage, halo_pitch, halo_a_bytes, halo_T, halo_cs;
    int halo_pair;       // 1: CTA pairs run M = 256 cta_group::2 MMAs (halo_cs == 2, no multicast)
    CUtensorMap tmap_b2;
    int halo_gtaps;
    int halo_a_off[4][16];
    int halo_out_py[4], halo_out_px[4];
    int halo_nsub;
    HaloGroup* d_groups;
    HaloTap* d_taps;
    int n_groups, n_taps;
    int dbg;             // RNR_CONV_DBG ablation bits (profiling only)
    BnFin bnf;           // fused BatchNorm finalize (rnr_conv_plan_set_bn; halo kernel only)
    GStats gst;          // fused BatchNorm-backward statistics (rnr_conv_plan_set_gstats; halo kernel only)
};

struct WgradParams {
    ViewD aviews[RNR_MAX_VIEWS];
    ViewD gviews[4];
    const rnr_wtap_t* taps;      // device
    int n_taps, a_dtype, g_dtype;
    int cout;
    int mN, mY, mX;
    float* dw;
    int64_t s_co, s_ci;
};

struct rnr_wgrad_plan {
    WgradParams p;
    int impl;
    rnr_wtap_t* d_taps;
    rnr_wtap_t* h_taps;
    // SIMT tiling
    int n_tiles;         // (tap, ci-block, co-block) output tiles
    int* d_tile_tab;     // [n_tiles,3] = tap, ci block, co block
    int splitk;
    // tcgen05 path
    CUtensorMap tmap_a[RNR_MAX_VIEWS];
    CUtensorMap tmap_g[4];
    int th, tw, tiles_y, tiles_x;
    int smem_bytes, grid, stages;
    int n_work;          // tc work items
    int vec;             // 1: dW destination is ci-contiguous -> 128-bit vector reductions
    int swap;            // 1: M = input channels, N = output channels
    int tc_pair;         // 1: CTA pairs (cta_group::2), 256 output channels per work item
    int tc_stage_bytes, tc_stages, tc_acc_cols;   // wgrad_tc operand ring / accumulator width (N = 256 work items need 96 KB stages)
    // halo-reuse / multi-tap kernel (wgrad_halo.cu)
    int halo, halo_pitch, halo_a_stride, halo_a_bytes, halo_stage_bytes, halo_nbuf;
    int* d_work_tab;     // [n_work, 8]
};

int rnr_encode_view_map(CUtensorMap* map, const rnr_view_t& v, int dtype, int box_c, int box_x, int box_y);
int rnr_conv_tc_prepare(rnr_conv_plan* plan, const rnr_conv_problem_t* prob);
int rnr_conv_tc_run(const rnr_conv_plan* plan, cudaStream_t stream);
int rnr_conv_halo_prepare(rnr_conv_plan* plan, const rnr_conv_problem_t* probs, int nsub);
int rnr_conv_halo_run(const rnr_conv_plan* plan, cudaStream_t stream);
int rnr_wgrad_tc_prepare(rnr_wgrad_plan* plan, const rnr_wgrad_problem_t* prob);
int rnr_wgrad_tc_run(const rnr_wgrad_plan* plan, cudaStream_t stream);
int rnr_wgrad_halo_prepare(rnr_wgrad_plan* plan, const rnr_wgrad_problem_t* prob);
int rnr_wgrad_halo_run(const rnr_wgrad_plan* plan, cudaStream_t stream);
