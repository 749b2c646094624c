/*
 * rnr_b200.h -- C ABI of librnr_b200.so, the B200 (sm_100a) implementation of the
 * per-view deferred-relighting hot path of LansburyCH/relightable-nr.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream and returns
 * a cudaError_t as int (0 == success).  No torch types cross this boundary.  Each
 * declaration cites the reference interface (file:line under /root/reference) that a
 * maintainer would re-bind to it; INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - "act"  : activations, fp16, channels-last, stored with a 1-pixel halo
 *              [N, H+2, W+2, C]  (reflect halo for forward tensors, zero halo for gradients)
 *   - "raw"  : pre-BatchNorm conv outputs, fp32 channels-last [N, H, W, C]
 *   - "grad" : gradients w.r.t. raw conv outputs, bf16, [N, H+2, W+2, C], zero halo
 *   - stream : cudaStream_t passed as void*
 */
#ifndef RNR_B200_H
#define RNR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* library info                                                                                */
/* ------------------------------------------------------------------------------------------ */
const char* rnr_version(void);
const char* rnr_last_error(void);          /* text of the last failing call on this thread */
int  rnr_device_sm_count(int device);
unsigned long long rnr_launch_count(void); /* kernels launched by this library so far (process-wide) */

/* ------------------------------------------------------------------------------------------ */
/* Generic implicit-GEMM convolution problem                                                   */
/*   replaces every ATen/cuDNN Conv2d / ConvTranspose2d (fwd, dgrad) of the U-Net:             */
/*   pytorch_prototyping/pytorch_prototyping.py:112-115,155-160,242-264 (and their autograd).  */
/*                                                                                            */
/*   out[n, y*my+py, x*mx+px, co] = epi( sum_j  A_j[n, y+dy_j, x+dx_j, c0_j .. c0_j+BK)         */
/*                                            . Wmat[co, j*BK .. (j+1)*BK) )                   */
/*   A_j is one of up to 8 strided 4-D "views" (C,X,Y,N); out-of-range coordinates read 0.     */
/* ------------------------------------------------------------------------------------------ */
#define RNR_MAX_VIEWS 8

enum { RNR_F16 = 0, RNR_BF16 = 1, RNR_F32 = 2 };

typedef struct {
    const void* ptr;       /* element (c=0,x=0,y=0,n=0) */
    int32_t dim[4];        /* extents  (C, X, Y, N) */
    int64_t stride[4];     /* element strides (C stride must be 1) */
} rnr_view_t;

typedef struct {           /* one K-step = BK consecutive channels of one view at one tap */
    int16_t view;
    int16_t c0;
    int16_t dx;
    int16_t dy;
} rnr_kstep_t;

enum {
    RNR_EPI_BIAS  = 1,     /* += bias[co]                                            */
    RNR_EPI_TANH  = 2,     /* tanh() after bias                                       */
    RNR_EPI_STATS = 4,     /* per-tile channel sums: stats[(tile_m*2+{0,1})*ldstats+co] */
    RNR_EPI_GSTATS = 8     /* data-gradient plan: reserve room for rnr_conv_plan_set_gstats (no effect until that call) */
};

typedef struct {
    /* A side */
    rnr_view_t views[RNR_MAX_VIEWS];
    int32_t    n_views;
    int32_t    ab_dtype;           /* RNR_F16 or RNR_BF16: dtype of views and Wmat */
    int32_t    bk;                 /* channels per K-step: 16, 32 or 64 */
    int32_t    n_ksteps;
    const rnr_kstep_t* ksteps;     /* HOST pointer, n_ksteps entries (copied at plan creation) */
    /* B side: Wmat [n_rows_w, n_ksteps*bk], K contiguous */
    const void* wmat;
    int32_t    n_rows_w;           /* rows of Wmat (>= cout, multiple of 16) */
    int32_t    cout;               /* valid output channels */
    /* M space */
    int32_t    mN, mY, mX;         /* extents of the (n,y,x) iteration space */
    int32_t    th, tw;             /* tile = th x tw pixels, th*tw == 128 */
    /* output */
    void*      out;
    int32_t    out_dtype;          /* RNR_F32 / RNR_F16 / RNR_BF16 */
    int64_t    out_sn, out_sy, out_sx;   /* element strides */
    int32_t    out_my, out_mx, out_py, out_px;
    int32_t    epi;                /* RNR_EPI_* flags */
    const float* bias;
    float*     stats;              /* [n_tiles_m, 2, ldstats] */
    int32_t    ldstats;
} rnr_conv_problem_t;

typedef struct rnr_conv_plan rnr_conv_plan_t;   /* opaque: device k-step table + TMA tensor maps */

/* impl: 0 = SIMT validation kernel, 1 = tcgen05/TMA kernel */
int  rnr_conv_plan_create(const rnr_conv_problem_t* prob, int impl, rnr_conv_plan_t** plan);
/* n <= 4 sub-problems differing only in tap offsets (dx, dy), wmat (stacked row-wise in one buffer: sub s at byte offset
 * s * n_rows_w * n_ksteps * bk * 2 from sub 0) and output parity (out_py, out_px) as ONE launch: the four output parities of
 * nn.ConvTranspose2d(4, 2, 1) forward (pytorch_prototyping.py:155-160) or of the data gradient of nn.Conv2d(4, stride 2)
 * (:242-264).  Returns cudaErrorNotSupported without a plan when they cannot be fused (create one plan per problem then). */
int  rnr_conv_plan_create_multi(const rnr_conv_problem_t* probs, int n, int impl, rnr_conv_plan_t** plan);
void rnr_conv_plan_destroy(rnr_conv_plan_t* plan);
int  rnr_conv_run(const rnr_conv_plan_t* plan, void* stream);
int  rnr_conv_plan_tiles_m(const rnr_conv_plan_t* plan);
/* BatchNorm-backward statistics of the PRODUCER layer(s) of a data-gradient plan's output, accumulated in the plan's epilogue.
 * The reference differentiates nn.BatchNorm2d + (Leaky)ReLU + Dropout2d (pytorch_prototyping.py:177-197, :250-272) with autograd:
 * dgamma = sum(gg * xhat), dbeta = sum(gg), gg = g * drop * lrelu'(.).  Both sums are linear in the incoming gradient g, so each
 * consumer layer's data-gradient launch adds its share (fp64 atomics into totals[2, C]: sum(gg), sum(gg * (raw - mean))) and
 * rnr_bn_bwd_apply_src finishes the layer in one pass.  Segment i covers output columns [c_lo, c_hi) (one concatenated input of
 * the consumer); raw == NULL leaves that segment alone.  H, W: interior size of the activation; pad = 1 when the plan's output grid
 * is the reflect-padded plane.  Returns cudaErrorNotSupported (no error text) when the plan cannot carry the statistics. */
typedef struct {
    const void*  raw;        /* producer's pre-BatchNorm conv output [N, H, W, C], 16-bit */
    int32_t      raw_dtype;
    int32_t      C;
    const float* scale;      /* gamma * invstd            [C] */
    const float* shift;      /* beta - mean * scale       [C] */
    const float* mean;       /*                           [C] */
    const float* drop;       /* Dropout2d scale [N, C] or NULL */
    float        slope;      /* LeakyReLU slope (0: ReLU) */
    int32_t      c_lo, c_hi;
    double*      totals;     /* [2, C] */
} rnr_gstat_seg_t;
int  rnr_conv_plan_set_gstats(rnr_conv_plan_t* plan, const rnr_gstat_seg_t* segs, int nseg, int H, int W, int pad);
/* profiling aid: device buffer [4 CTAs][4 roles][64] of clock64 stamps written by the halo kernel (NULL = off) */
int  rnr_debug_set_trace(long long* buf);
/* rows of `stats` the plan writes ([rows, 2, ldstats]; the caller sizes / zero-fills the buffer accordingly) */
int  rnr_conv_plan_stat_rows(const rnr_conv_plan_t* plan);
/* Fuse nn.BatchNorm2d's batch-statistics finalize (pytorch_prototyping.py:177-197: mean / biased variance of the conv output,
 * running_mean / running_var / num_batches_tracked update) into the plan's kernel: its last CTA reduces the per-CTA partial
 * sums and writes mean / invstd / scale = gamma*invstd / shift = beta - mean*scale.  cudaErrorNotSupported (no error text)
 * when the plan does not run on the halo kernel with RNR_EPI_STATS -- the caller then launches rnr_bn_finalize itself.
 * `ticket`: one zero-initialised int per plan.  running_* / num_batches_tracked may be NULL (no update). */
int  rnr_conv_plan_set_bn(rnr_conv_plan_t* plan, const float* gamma, const float* beta, double count, float eps, float momentum,
                          float* mean, float* invstd, float* scale, float* shift, float* running_mean, float* running_var,
                          long long* num_batches_tracked, int* ticket, int enabled);

/* ------------------------------------------------------------------------------------------ */
/* Weight-gradient problem (autograd of Conv2d/ConvTranspose2d w.r.t. weight)                  */
/*   dW[co*s_co + ci*s_ci + off_t] (+)= sum_{n,y,x} G[n,y,x,co] * A_t[n, y+dy_t, x+dx_t, ci]    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int16_t view;        /* A view of this tap */
    int16_t dx, dy;
    int16_t gview;       /* G view of this tap (parity classes of ConvTranspose use different G views) */
    int32_t c0;          /* first input channel (within the A view) */
    int32_t ci0;         /* first ci index in dW for this entry */
    int32_t nci;         /* number of channels */
    int64_t off;         /* element offset of this tap in dW */
} rnr_wtap_t;

typedef struct {
    rnr_view_t aviews[RNR_MAX_VIEWS];
    rnr_view_t gviews[4];
    int32_t    n_aviews, n_gviews;
    int32_t    a_dtype, g_dtype;
    int32_t    n_taps;
    const rnr_wtap_t* taps;        /* HOST pointer */
    int32_t    cout;
    int32_t    mN, mY, mX;         /* pixel iteration space (shared by G and A views) */
    float*     dw;                 /* fp32, accumulated with atomics: must be zeroed by caller */
    int64_t    s_co, s_ci;
} rnr_wgrad_problem_t;

typedef struct rnr_wgrad_plan rnr_wgrad_plan_t;
int  rnr_wgrad_plan_create(const rnr_wgrad_problem_t* prob, int impl, rnr_wgrad_plan_t** plan);
void rnr_wgrad_plan_destroy(rnr_wgrad_plan_t* plan);
int  rnr_wgrad_run(const rnr_wgrad_plan_t* plan, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Weight preparation: fp32 master weights -> 16-bit K-major GEMM matrices                     */
/*   dst[r*ld + t*cpad + c] = src[r*s_r + c*s_c + tapoff[t]]   (0 for c >= nc or r >= nr)       */
/*   chunked: dst[r*ld + ((c/64)*ntaps + t)*64 + c%64]  (taps of a 64-channel chunk adjacent)    */
/* ------------------------------------------------------------------------------------------ */
int rnr_weight_prep(const float* src, void* dst, int dst_dtype,
                    int nr, int nr_pad, int nc, int cpad, int ntaps,
                    int64_t s_r, int64_t s_c, const int32_t* tapoff_dev, int chunked, void* stream);

/* batched form: all layers' matrices in ONE launch (the job table lives in a plan) */
typedef struct {
    const float* src;          /* fp32 parameter (element offset already applied) */
    void*        dst;          /* 16-bit matrix [nr_pad, ntaps*cpad] */
    int32_t      dst_dtype;    /* RNR_F16 / RNR_BF16 */
    int32_t      nr, nr_pad, nc, cpad, ntaps;
    int64_t      s_r, s_c;     /* element strides of row / column index in src; min(s_r, s_c) = taps per (row, col) pair <= 16 */
    int32_t      tapoff[16];   /* source tap index of each destination tap */
    int32_t      chunked;      /* 0: columns [tap][cpad]; 1: columns [cpad/64][tap][64] (all taps of a 64-channel chunk adjacent) */
    int64_t      ld;           /* destination row pitch in elements; 0 = ntaps*cpad (a job may fill a column block of a wider matrix) */
} rnr_wprep_job_t;
typedef struct rnr_wprep_plan rnr_wprep_plan_t;
int  rnr_wprep_plan_create(const rnr_wprep_job_t* jobs, int njobs, rnr_wprep_plan_t** plan);
void rnr_wprep_plan_destroy(rnr_wprep_plan_t* plan);
int  rnr_wprep_run(const rnr_wprep_plan_t* plan, void* stream);

/* Weight-gradient un-transpose (one launch for all layers): the weight-gradient problem may point `dw` at a scratch buffer in
 * GEMM order [tap][co][ci] (s_co = cin, s_ci = 1, tap offset = tap*cout*cin), which lets the tcgen05 kernel reduce with 128-bit
 * vector atomics; this pass writes the parameter's own layout (autograd of nn.Conv2d / nn.ConvTranspose2d w.r.t. weight,
 * pytorch_prototyping.py:112-115,155-160,242-264):   dst[co*s_co + ci*s_ci + t] = src[(t*cout + co)*cin + ci]               */
typedef struct {
    const float* src;          /* scratch [ntaps, cout, cin] */
    float*       dst;          /* parameter-layout gradient */
    int32_t      cout, cin, ntaps;
    int64_t      s_co, s_ci;   /* element strides of co / ci in dst (taps are contiguous) */
} rnr_wunpack_job_t;
typedef struct rnr_wunpack_plan rnr_wunpack_plan_t;
int  rnr_wgrad_unpack_plan_create(const rnr_wunpack_job_t* jobs, int njobs, rnr_wunpack_plan_t** plan);
void rnr_wgrad_unpack_plan_destroy(rnr_wunpack_plan_t* plan);
int  rnr_wgrad_unpack_run(const rnr_wunpack_plan_t* plan, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Optimiser: torch.optim.Adam(lr) of train_rnr.py:376,618-623 / train_dnr.py:176,259-262       */
/*   (fused-Adam arithmetic: m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;                        */
/*    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)), step counter resident on the device   */
/* ------------------------------------------------------------------------------------------ */
/* A 16-bit GEMM matrix family re-derived from an updated conv weight inside the optimiser pass (what rnr_wprep_run produces):
 * `nsub` sub-matrices of `sub_rows` rows stacked row-wise behind `base` (the output-parity sub-problems of a stride-2 layer),
 * columns chunk-major [64-channel chunk][dst tap][64].  inv[t] = s*16 + k places source tap t (kh*KW + kw) at tap slot k of
 * sub-matrix s (-1: unused).  Forward family: row = output channel, column channel = input channel.  Data-gradient family:
 * row = input channel - r0 for r0 <= ci < r1, column channel = output channel.  base = NULL: family not produced here. */
typedef struct {
    void*        base;
    int64_t      ld;           /* row pitch in elements */
    int32_t      dtype;        /* RNR_F16 / RNR_BF16 */
    int32_t      ntaps;        /* tap slots per sub-matrix */
    int32_t      sub_rows;
    int32_t      r0, r1;       /* data-gradient family only */
    int8_t       inv[16];
} rnr_wmat_t;
typedef struct {               /* a conv weight whose gradient sits in GEMM order (see rnr_wunpack_job_t) */
    float*       scratch;      /* [ntaps, cout, cin] gradient; re-zeroed by the pass when zero_grad */
    float*       p;            /* fp32 master weight, parameter layout: (co, ci, t) at co*s_co + ci*s_ci + t */
    float*       m;            /* exp_avg,    parameter layout */
    float*       v;            /* exp_avg_sq, parameter layout */
    float*       gdst;         /* optional: also write the un-transposed gradient here (NULL: never materialised) */
    int32_t      cout, cin, ntaps;
    int64_t      s_co, s_ci;
    rnr_wmat_t   fwd, dgrad;   /* next step's GEMM matrices, written from the updated weight in the same pass (base NULL: skip) */
} rnr_adam_wjob_t;
typedef struct {               /* a plain tensor (bias, BatchNorm affine, texture level, SH coefficients) */
    float*       p;
    float*       g;            /* gradient; re-zeroed by the pass when zero_grad */
    float*       m;
    float*       v;
    int64_t      n;
} rnr_adam_job_t;
typedef struct rnr_adam_plan rnr_adam_plan_t;
int  rnr_adam_plan_create(const rnr_adam_wjob_t* wjobs, int n_wjobs, const rnr_adam_job_t* jobs, int n_jobs, rnr_adam_plan_t** plan);
void rnr_adam_plan_destroy(rnr_adam_plan_t* plan);
/* One Adam step over the plan's tensors (<= 2 launches).  `step`: device float, number of steps taken so far; with
 * advance_step the last block of the tensor-job launch increments it.  gscale multiplies every gradient (1/world). */
int  rnr_adam_run(const rnr_adam_plan_t* plan, float* step, float lr, float beta1, float beta2, float eps, float gscale,
                  int zero_grad, int advance_step, void* stream);
/* out[0] = sums[2]/cnt + sums[0]/sums[1]/R*w_chrom + sum(extra[0..n_extra)): the scalar of train_rnr.py:608 from the device
 * accumulators of rnr_tail_fwd and the small-loss kernels (no host round trip, no ATen glue) */
int  rnr_loss_combine(double* sums, double cnt, double R, double w_chrom, const double* extra, int n_extra, float* out,
                      int n_clear, void* stream);   /* then zeroes sums[0..n_clear): the accumulators clean up behind themselves */
/* nn.Dropout2d channel masks (pytorch_prototyping.py:181,190,255,268: p = 0.1): out[i] = 0 w.p. p else 1/(1-p); counter-based
 * generator keyed by (seed, *counter); *counter is advanced by the launch (fresh masks on every CUDA-graph replay) */
int  rnr_dropout_masks(float* out, int n, float p, unsigned long long seed, unsigned long long* counter, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* BatchNorm2d (batch statistics) + activation + Dropout2d                                     */
/*   replaces nn.BatchNorm2d / LeakyReLU / ReLU / Dropout2d of pytorch_prototyping.py:177-197,  */
/*   250-272, 471-476 (forward and backward)                                                   */
/* ------------------------------------------------------------------------------------------ */
/* partial sums [T,2,ld] -> mean/invstd/scale/shift (+ running stats update, momentum)         */
int rnr_bn_finalize(const float* partials, int T, int ld, int C, double count,
                    const float* gamma, const float* beta, float eps,
                    float* mean, float* invstd, float* scale, float* shift,
                    float* running_mean, float* running_var, float momentum, void* stream);

/* act = drop * act(raw*scale + shift), written fp16 with reflect halo [N,H+2,W+2,C];
 * act_bf16 (optional, same layout) receives a bf16 copy: the B operand of the weight-gradient MMA
 * (kind::f16 needs both operands in the same 16-bit format and gradients are bf16)               */
/* nn.BatchNorm2d batch statistics WITHOUT a finalize launch: rnr_conv_plan_set_stat_totals makes the conv plan add its per-CTA
 * sums to totals [2, C] (fp64, zero before the first launch); rnr_bn_act_fwd_tot derives mean / biased variance / invstd / scale /
 * shift from them in every block (arithmetic of rnr_bn_finalize), applies scale / shift + (Leaky)ReLU + Dropout2d like
 * rnr_bn_act_fwd, publishes mean / invstd / scale / shift and the running-statistics update from its first block, and re-zeroes
 * totals / re-arms ticket (int32, zero before the first call) from the last block that read them. */
int rnr_conv_plan_set_stat_totals(rnr_conv_plan_t* plan, double* totals);
int rnr_bn_act_fwd_tot(const void* raw, int raw_dtype, double* totals, int* ticket, double count, const float* gamma,
                       const float* beta, float eps, float* mean, float* invstd, float* scale, float* shift, float* running_mean,
                       float* running_var, float momentum, const float* drop, float slope, void* act, void* act_bf16, int N, int H,
                       int W, int C, void* stream);
int rnr_bn_act_fwd(const void* raw, int raw_dtype /* RNR_F32 or 16-bit */, const float* scale, const float* shift,
                   const float* drop /* [N,C] or NULL */, float slope,
                   void* act, void* act_bf16, int N, int H, int W, int C, void* stream);

typedef struct {
    const void* ptr;      /* bf16 or fp32 */
    int32_t dtype;
    int32_t fold;         /* 1: tensor is [N,H+2,W+2,ld] = grad w.r.t. the reflect-padded input: fold halo */
    int32_t ld;           /* channel pitch */
    int32_t c0;           /* first channel */
} rnr_gsrc_t;

/* pass 1: gz = (sum of sources) * drop * act'(raw*scale+shift); writes gz (bf16, zero-halo layout)
 *         and per-block partial sums of gz and gz*xhat: partials [T,2,C]                        */
int rnr_bn_bwd_reduce(const rnr_gsrc_t* srcs, int nsrc, const float* raw,
                      const float* scale, const float* shift, const float* mean, const float* invstd,
                      const float* drop, float slope,
                      void* gz, float* partials, int* T_out,
                      int N, int H, int W, int C, void* stream);
/* finalize: dgamma, dbeta, c1 = mean(gz), c2 = mean(gz*xhat) and (optional) the fused per-channel coefficients
 * coef [3,C] = (A, B, D) with  gamma*invstd*(gz - c1 - xhat*c2) == A*gz + B*raw + D                      */
int rnr_bn_bwd_finalize(const float* partials, int T, int C, double count,
                        float* dgamma, float* dbeta, float* c1, float* c2,
                        const float* gamma, const float* mean, const float* invstd, float* coef, void* stream);
/* pass 1 + finalize in ONE launch: as rnr_bn_bwd_reduce, but the block partial sums are added (fp64 atomics) into
 * totals [2,C] (double, must be 0 before the first call; the kernel re-zeroes it) and the last block to finish (ticket: an
 * int32 in device memory, 0 before the first call, re-armed by the kernel) writes dbeta = sum(gz), dgamma = sum(gz*xhat)
 * (either may be NULL) and, when coef != NULL, the coefficients (A, B, D) of rnr_bn_bwd_apply.                          */
/* BatchNorm backward from totals that rnr_conv_plan_set_gstats launches accumulated: gz = A*gg + B*raw + D in ONE pass over the
 * gradient sources (no reduction pass); writes dgamma / dbeta, re-zeroes totals, re-arms ticket.  C = 8 x a power of two. */
int rnr_bn_bwd_apply_src(const rnr_gsrc_t* srcs, int nsrc, const void* raw, int raw_dtype, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* gamma, const float* drop, float slope, void* gz,
                         double* totals, int* ticket, double count, float* dgamma, float* dbeta, int N, int H, int W, int C,
                         void* stream);
int rnr_bn_bwd_reduce_fin(const rnr_gsrc_t* srcs, int nsrc, const void* raw, int raw_dtype,
                          const float* scale, const float* shift, const float* mean, const float* invstd,
                          const float* drop, float slope, void* gz, double* totals, int* ticket, double count,
                          float* dgamma, float* dbeta, const float* gamma, float* coef,
                          int N, int H, int W, int C, void* stream);
/* pass 2 (in place): gz <- A*gz + B*raw + D */
int rnr_bn_bwd_apply(void* gz, const void* raw, int raw_dtype, const float* coef, int N, int H, int W, int C, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* layout glue of the module-level API (NCHW fp32 <-> channels-last 16-bit)                    */
/*   network.RenderingNet.forward(input [N,C,H,W]) -> [N,Cout,H,W]   network.py:251-253        */
/* ------------------------------------------------------------------------------------------ */
int rnr_pack_nchw_to_act(const float* src, void* act, void* act_bf16 /* optional */, int N, int C, int Cpad, int H, int W, void* stream);
int rnr_unpack_nhwc_to_nchw(const float* src, float* dst, int N, int C, int ld, int H, int W, void* stream);
/* grad_out NCHW fp32 (d/d tanh-output) -> gz = grad*(1-t^2) bf16 zero-halo, + per-block bias partial sums */
int rnr_tanh_bwd_pack(const float* grad_nchw, const float* tanh_nhwc, void* gz, float* dbias /* [C] atomics, pre-zeroed */,
                      int N, int C, int ld, int H, int W, void* stream);
/* folded grad w.r.t. reflect-padded input [N,H+2,W+2,ld] -> NCHW fp32 [N,C,H,W] (channels c0..c0+C) */
int rnr_fold_to_nchw_add(const void* gpad, int dtype, float* dst, int N, int C, int c0, int ld, int H, int W,
                         const float* add, int nadd, void* stream);
int rnr_fold_to_nchw(const void* gpad, int dtype, float* dst, int N, int C, int c0, int ld, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Neural texture: network.TextureMapper.forward (network.py:67-91), misc.interpolate_bilinear */
/* (misc.py:5-42), TextureMapper.flatten_mipmap (network.py:93-99), network.Interpolater       */
/* (network.py:322-337).  tex / gtex / sizes are HOST arrays of L device pointers / sizes.       */
/*   uv [N,H,W,2], sh [N,H,W,9] or NULL, out / gout NCHW fp32 [N,C,H,W]; textures [S,S,C] fp32.  */
/* ------------------------------------------------------------------------------------------ */
int rnr_texmap_fwd(const float* const* tex, const int* sizes, int L, int C, const float* uv, const float* sh,
                   int sh_start, float* out_nchw, int N, int H, int W, void* stream);
int rnr_texmap_bwd(float* const* gtex /* accumulated */, const int* sizes, int L, int C, const float* uv, const float* sh,
                   int sh_start, const float* gout_nchw, int N, int H, int W, void* stream);
int rnr_flatten_mipmap(const float* const* tex, float* const* gtex, const int* sizes, int L, int C, int c0, int nc,
                       float* out /* [S0,S0,nc] */, const float* gout, int backward, void* stream);
/* data [Nd(1|N),Hd,Wd,C]; xs, ys [N,M] -> out [N,M,C]; bwd accumulates into gdata */
int rnr_bilinear_fwd(const float* data, int Nd, int Hd, int Wd, int C, const float* xs, const float* ys,
                     float* out, int64_t M, int N, void* stream);
int rnr_bilinear_bwd(float* gdata, int Nd, int Hd, int Wd, int C, const float* xs, const float* ys,
                     const float* gout, int64_t M, int N, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Rays: network.RaySampler.forward (network.py:445-472), network.RayRenderer.forward           */
/* (network.py:481-527) and its backward.  tbn [P,3,3], vdt [P,3], alpha [P], pivots [3,R];      */
/* rays_dir [P,3,R], rays_uv [P,2,R]; rays_lt / rays_color [N,R,3,H,W]; images NCHW.             */
/* ------------------------------------------------------------------------------------------ */
int rnr_ray_sampler_fwd(const float* tbn, const float* vdt, const float* alpha, const float* pivots, int R,
                        int reflect, float* rays_dir, float* rays_uv, float* rays_dir_tangent, int64_t P, void* stream);
int rnr_ray_render_fwd(const float* alb_s, const float* alb_d, const float* rays_uv, const float* rays_lt,
                       const float* lp, int Nl, int Hl, int Wl, int R, int Rd, int no_albedo, int separate,
                       float* out, float* out_s, float* out_d, float* ltt_s, float* ltt_d, float* rays_color,
                       int N, int H, int W, void* stream);
int rnr_ray_render_bwd(const float* alb_s, const float* alb_d, const float* rays_uv, const float* rays_lt,
                       const float* lp, int Nl, int Hl, int Wl, int R, int Rd, int no_albedo, int separate,
                       const float* g_out, const float* g_os, const float* g_od, const float* g_ls, const float* g_ld,
                       const float* ltt_s, const float* ltt_d, float* g_alb_s, float* g_alb_d, float* g_lt,
                       float* g_lp /* accumulated */, int N, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Per-pixel maps derived from the G-buffer: camera.get_view_dir_map (camera.py:5-32),          */
/* render.get_TBN_map (render.py:124-168), render.interp_vertex_attr (render.py:11-28)          */
/* ------------------------------------------------------------------------------------------ */
/* proj_inv, R_inv [N,3,3] -> view_dir_map [N,H,W,3] (world), optional camera-space map          */
int rnr_view_dir_map(const float* proj_inv, const float* R_inv, float* out_world, float* out_cam, int N, int H, int W,
                     void* stream);
/* faces_v [nf,3,3], faces_vt [nf,3,2] -> normalised per-face tangent [nf,3]; nan_flag[0] |= 1 on NaN */
int rnr_face_tangents(const float* faces_v, const float* faces_vt, float* tangent, int* nan_flag, int nf, void* stream);
/* normal_map [P,3], face_index_map [P] (i32, -1 = background) -> TBN [P,3,3] (columns T,B,N)     */
int rnr_tbn_map(const float* normal_map, const int* face_index_map, const float* tangent, float* tbn, int* nan_flag,
                int64_t P, int nf, void* stream);
/* attr [1|N,nv,A], faces [N,nf,3] i32, face_index_map [N,P] i32, weight_map [N,P,3] -> out [N,P,A] */
int rnr_interp_vertex_attr(const float* attr, int attr_batch, int nv, int A, const int* faces, int nf,
                           const int* face_index_map, const float* weight_map, float* out, int N, int64_t P, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Proxy-mesh rasterizer.  Replaces neural_renderer's projection (projection.py:6-53),          */
/* vertices_to_faces (vertices_to_faces.py:4-25), rasterize_cuda.forward_face_index_map         */
/* (cuda/rasterize_cuda.cpp:70-95, cuda/rasterize_cuda_kernel.cu:24-169), the vertical flip of   */
/* rasterize_rgbad (rasterize.py:313-321) and the attribute interpolation of                    */
/* network.Rasterizer.forward (network.py:176-214).                                             */
/* ------------------------------------------------------------------------------------------ */
/* vertices [1|N,nv,3], K/R [N,3,3], t [N,3], dist_coeffs [N,5]|NULL, offset/scale [N,2]|NULL -> uvz [N,nv,3] */
int rnr_project_vertices(const float* vertices, int v_batch, int nv, const float* K, const float* R, const float* t,
                         const float* dist_coeffs, const float* offset, const float* scale, float orig_size, float eps,
                         float* out_uvz, int N, void* stream);
/* per-face set-up.  Either faces_in [N,nf,3,3] is given (the reference extension's calling convention) or the faces are
 * gathered from uvz [N,nv,3] through faces_idx [1|N,nf,3] (and optionally written to faces_out).
 * Outputs: faces_inv [N,nf,3,3] (zeros for back faces, like torch.zeros_like in rasterize.py:163) and
 * bbox [N,nf,2] int32 = packed pixel bounding boxes (x0 | x1<<16, y0 | y1<<16; x0 > x1 = culled).           */
int rnr_raster_face_setup(const float* uvz, int nv, const int32_t* faces_idx, int f_batch, const float* faces_in, int nf,
                          int image_size, float* faces_out, float* faces_inv, int32_t* bbox, int N, void* stream);
typedef struct {               /* optional fused outputs of network.Rasterizer.forward; NULL members are skipped */
    const float* v;  const int32_t* f_v_idx;      /* [nv,3],  [nf,3] */
    const float* vt; const int32_t* f_vt_idx;     /* [nvt,2], [nf,3] */
    const float* vn; const int32_t* f_vn_idx;     /* [nvn,3], [nf,3] */
    const float* pose_R;                          /* [N,3,3] */
    const float* pose_t;                          /* [N,3]   */
    float* weight_pc;                             /* [N,is,is,3] perspective-correct weights (network.py:176-180) */
    float* uv_map;                                /* [N,is,is,2] wrapped to [0,1) (:187-190) */
    float* normal_map;                            /* [N,is,is,3] world, normalised (:197-200) */
    float* normal_map_cam;                        /* (:203-205) */
    float* position_map;                          /* (:208-210) */
    float* position_map_cam;                      /* (:213-214) */
} rnr_raster_attrs_t;
/* z-buffered coverage of every pixel; every output pixel is written (background: index -1, weights 0, depth `far`),
 * so no pre-fill is needed.  flip_y folds rasterize_rgbad's vertical flip into the store address.          */
int rnr_raster_tiles(const float* faces, const float* faces_inv, const int32_t* bbox, int nf, int image_size, float near,
                     float far, int flip_y, int32_t* face_index_map, float* weight_map, float* depth_map, float* alpha_map,
                     float* face_inv_map, const rnr_raster_attrs_t* attrs, int N, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Spherical harmonics: sph_harm.evaluate_sh_basis(lmax=2) (sph_harm.py:41-71),                 */
/* sph_harm.reconstruct_sh (:91-102) fwd/bwd, sph_harm.fit_sh_coeff (:74-88)                    */
/* ------------------------------------------------------------------------------------------ */
int rnr_sh_basis_l2(const float* dirs /* [P,3] */, float* out /* [P,9] */, int64_t P, void* stream);
/* general degree, fp64 table [P,(lmax+1)^2] (set-up time: network.py:557,581,696) */
int rnr_sh_basis(const float* dirs /* [P,3] */, double* out, int64_t P, int lmax, void* stream);
/* out[l,p,c] = sum_b basis[p,b] coeff[l,b,c] */
int rnr_sh_reconstruct(const float* basis, const float* coeff, float* out, int64_t P, int B, int Cc, int Lc, void* stream);
/* res[l,b,c] += scale * sum_p basis[p,b] v[l,p,c]   (res pre-zeroed / accumulated) */
int rnr_sh_project(const float* basis, const float* v, float* res, int64_t P, int B, int Cc, int Lc, float scale, void* stream);
/* pitched variants for one lighting: out rows ldo >= Cc apart (extra columns untouched) / v rows ldv >= Cc apart, optionally cleared
 * as they are consumed (the [P,4] envmap texels and their gradient accumulator of the fused step) */
int rnr_sh_reconstruct_ld(const float* basis, const float* coeff, float* out, int64_t P, int B, int Cc, int ldo, void* stream);
int rnr_sh_project_ld(const float* basis, float* v, float* res, int64_t P, int B, int Cc, int ldv, float scale, int zero_v, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Losses and optimiser: network.RaysLTChromLoss (network.py:395-411), masked cropped L1         */
/* (train_rnr.py:565-585), torch.optim.Adam step (train_rnr.py:376,622)                          */
/* ------------------------------------------------------------------------------------------ */
/* sums[0] += sum(diff), sums[1] += sum(alpha) (double, pre-zeroed); optional full outputs */
int rnr_chrom_loss_fwd(const float* rays_lt, const float* alpha, const float* img, int R, int N, int H, int W,
                       float* chrom, float* chrom_mean, float* diff, double* sums, void* stream);
int rnr_chrom_loss_bwd(const float* rays_lt, const float* alpha, const float* img, int R, int N, int H, int W,
                       const double* sums, const float* gscale_dev, float gscale, float* g_lt, int accumulate, void* stream);
/* loss_sum += weight * mean|out*a - gt*a| over the central crop; g_out = d(weight*loss)/d out */
int rnr_l1_masked(const float* out, const float* gt, const float* alpha, int N, int C, int H, int W, int crop,
                  float weight, float* g_out, double* loss_sum, void* stream);
int rnr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, int step, float gscale, void* stream);
/* Lighting L1 (train_rnr.py:571-579) for ONE lighting: est = basis[S,B] coeff[B,3]; *loss += sum_s |l_init - est| w_s with
 * w_s = w_cov (mask[s] != 0) or w_unc -- the caller passes weight / sample count; sgn[S,3] = d loss / d est (scatter it into
 * the coefficient gradient with rnr_sh_project). */
int rnr_lighting_l1(const float* basis, const float* coeff, const float* l_init, const unsigned char* mask, int S, int B,
                    float w_cov, float w_unc, float* sgn, double* loss, void* stream);
/* Albedo-mean loss (train_rnr.py:596-607) on tex6 = flatten_mipmap(channels 0:6) [P,6] against the initial flattened texture:
 * sums[8] (zeroed by the caller) receive the touched-texel counts / channel sums, gout[P,6] = d (w_alb*loss) / d tex6,
 * *loss += w_alb * loss (scatter gout into the levels with rnr_flatten_mipmap(backward = 1)). */
int rnr_albedo_mean_loss(const float* tex6, const float* init6, int64_t P, float w_alb, double* sums, float* gout, double* loss,
                         void* stream);

/* ------------------------------------------------------------------------------------------ */
/* gcn_lib/dense EdgeConv (network.DenseDeepGCN, network.py:256-315)                           */
/*   EdgeConv4D torch_vertex.py:23-35 over BasicConv = Conv2d(1x1) -> act -> BatchNorm2d        */
/*   torch_nn.py:55-64.  pq [V, 2C] = (P | Q): P = X (W1-W2)^T + b, Q = X W2^T (one GEMM by the  */
/*   caller; W = [W1 | W2] acts on cat[x_i, x_j - x_i]).  nbr [V,K] int32 = neighbour indices.   */
/* ------------------------------------------------------------------------------------------ */
/* a_k = act(P[v,c] + Q[nbr[v,k],c]); amax/amin [V,C] = max/min over k; sums[2C] (double, pre-zeroed, may be NULL)
 * += sum a_k, sum a_k^2 over all V*K edges (BatchNorm batch statistics)                                              */
int rnr_edgeconv_reduce(const float* pq, const int32_t* nbr, int V, int K, int C, float slope, float* amax, float* amin,
                        double* sums, void* stream);
/* out[v,c] = BN(max_k a_k) (+ residual): the affine BN map commutes with max (scale >= 0) or turns it into min (scale < 0);
 * gamma == NULL: no normalisation.  training: batch statistics from sums / count (+ running-stat update), else running stats */
int rnr_edgeconv_finish(const float* amax, const float* amin, const double* sums, double count, const float* gamma,
                        const float* beta, float eps, float* running_mean, float* running_var, float momentum, int training,
                        const float* residual, float* out, int V, int C, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused producers / consumers around the U-Net (csrc/fused.cu): the same operators as above,  */
/* composed so that no intermediate crosses HBM in a layout its consumer cannot use directly.  */
/* ------------------------------------------------------------------------------------------ */
/* network.TextureMapper.forward (network.py:67-91) + RaySampler.forward x2 (network.py:445-472, 'reflect' with pivots_s
 * [3,Rs] then 'diffuse' with pivots_d [3,Rd]) + the input assembly torch.cat((rays_dir, normal, view_dir, neural_img), 1)
 * (train_rnr.py:530-533) -> `act`: the first convolution's operand, fp16 [N,H+2,W+2,Cpad] channels-last with reflect halo
 * (channels r*3+c | normal | view_dir | texture | zero padding), optional bf16 copy, rays_uv [N,H,W,2,Rs+Rd], and
 * albedo [N,H,W,8] = texture channels 0..7 (diffuse 0..2, specular 3..5, train_rnr.py:515-516).
 * Rs = Rd = 0 with normal = view_dir = NULL gives the DNR input (train_dnr.py:252).                                   */
int rnr_head_fwd(const float* const* textures, const int* sizes, int n_levels, int C, const float* uv, const float* sh,
                 int sh_start, const float* tbn, const float* view_dir_tangent, const float* alpha, const float* normal,
                 const float* view_dir, const float* pivots_s, int Rs, const float* pivots_d, int Rd, void* act,
                 void* act_bf16, int Cpad, float* rays_uv, float* albedo, int N, int H, int W, void* stream);
/* rays_lt = (raw*0.5+0.5)*2 (train_rnr.py:535-536; raw = tanh output of the last convolution, NHWC with pitch ldraw),
 * RayRenderer.forward(seperate_albedo=True) (network.py:481-527) -> final [N,3,H,W]; RaysLTChromLoss (network.py:395-411)
 * and the cropped alpha-masked L1 (train_rnr.py:565-585) as sums (double[3], pre-zeroed): sum(diff), sum(alpha),
 * sum|final*a - gt*a|.  aux [N,H,W,12] keeps what the backward re-uses; lp4 [Hl*Wl,4] = envmap texels (r,g,b,unused).                                             */
int rnr_tail_fwd(const float* raw, int ldraw, const float* rays_uv, const float* albedo, const float* lp4, int Hl, int Wl,
                 const float* alpha, const float* img_gt, int Rs, int Rd, int N, int H, int W, int crop, float* final_img,
                 float* aux, double* sums, void* stream);
/* backward of rnr_tail_fwd for loss = w_l1*L1 + w_chrom*chrom: gz (bf16 [N,H+2,W+2,ldg], zero halo) = d loss / d (pre-tanh
 * output), dbias[3R] += its per-channel sums (bias gradient of the last convolution), g_alb [N,6,H,W] = d loss / d texture
 * channels 0..5, g_lp4 [Hl*Wl,4] += d loss / d envmap (rgb + one unused lane: 128-bit vector reductions).               */
int rnr_tail_bwd(const float* raw, int ldraw, const float* rays_uv, const float* albedo, const float* lp4, int Hl, int Wl,
                 const float* alpha, const float* img_gt, int Rs, int Rd, int N, int H, int W, int crop, const float* aux,
                 const double* sums, float w_l1, float w_chrom, void* gz, int ldg, float* dbias, float* g_alb,
                 float* g_lp4, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* neural_renderer.cuda: the cold entry points (never on the relighting path; provided so that  */
/* the extension's seven functions all exist).  Same arithmetic as the reference kernels.        */
/* ------------------------------------------------------------------------------------------ */
/* rasterize_cuda.cpp:97-122 / rasterize_cuda_kernel.cu:172-242: rgb_map [B,is,is,3] (+ 8 sampling indices / weights per pixel) from
 * the per-face texture cubes textures [B,nf,ts,ts,ts,3]; background pixels are left untouched */
int rnr_nr_forward_texture_sampling(const float* faces, const float* textures, const int32_t* face_index_map, const float* weight_map,
                                    const float* depth_map, float* rgb_map, int32_t* sampling_index_map, float* sampling_weight_map,
                                    int batch, int num_faces, int image_size, int texture_size, float eps, void* stream);
/* rasterize_cuda.cpp:124-148 / kernel :245-505: silhouette gradient into grad_faces [B,nf,3,3] (front faces overwritten, back faces untouched) */
int rnr_nr_backward_pixel_map(const float* faces, const int32_t* face_index_map, const float* rgb_map, const float* alpha_map,
                              const float* grad_rgb_map, const float* grad_alpha_map, float* grad_faces, int batch, int num_faces,
                              int image_size, float eps, int return_rgb, int return_alpha, void* stream);
/* rasterize_cuda.cpp:150-167 / kernel :507-541: grad_textures += scatter of grad_rgb_map through the sampling indices / weights */
int rnr_nr_backward_textures(const int32_t* face_index_map, const float* sampling_weight_map, const int32_t* sampling_index_map,
                             const float* grad_rgb_map, float* grad_textures, int batch, int num_faces, int image_size, int texture_size,
                             void* stream);
/* rasterize_cuda.cpp:169-191 / kernel :543-591: grad_faces += d depth / d vertices */
int rnr_nr_backward_depth_map(const float* faces, const float* depth_map, const int32_t* face_index_map, const float* face_inv_map,
                              const float* weight_map, const float* grad_depth_map, float* grad_faces, int batch, int num_faces,
                              int image_size, void* stream);
/* load_textures_cuda.cpp:20-39 / kernel :25-121: texture cubes [nf,ts,ts,ts,3] of the faces with is_update != 0 from image [H,W,3] through
 * the uv triangles faces [nf,3,2] (wrapped in place: 0 REPEAT, 1 MIRRORED_REPEAT, 2 CLAMP_TO_EDGE, 3 CLAMP_TO_BORDER) */
int rnr_nr_load_textures(const float* image, float* faces, float* textures, const int32_t* is_update, int num_faces, int texture_size,
                         int image_height, int image_width, int texture_wrapping, int use_bilinear, void* stream);
/* create_texture_image_cuda.cpp:18-33 / kernel :9-117: tiled atlas image [th*tso, tw*tso, 3] from the texture cubes */
int rnr_nr_create_texture_image(const float* vertices_all, const float* textures, float* image, int64_t image_numel, int num_faces,
                                int texture_size_in, int texture_size_out, int tile_width, float eps, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Validation metrics on the device: metric.py:19-84 (masked MAE / MSE / PSNR inputs)           */
/*   box [N,4] int32 pre-set to {W,-1,H,-1}, cnt [N] u64 and sums [N,4] f64 pre-zeroed:          */
/*   box = bounding box of mask == 1, cnt = its pixel count, sums = {sum|d|, sum d^2} over the   */
/*   image and over the box with d = (est - gt) where mask == 1, 0 elsewhere                     */
/* ------------------------------------------------------------------------------------------ */
int rnr_metric_sums(const float* est, const float* gt, const float* mask, int N, int C, int H, int W, int* box,
                    unsigned long long* cnt, double* sums, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Light-probe stitching: stitch_lp.py:22-35 (camera2ray, spherical_mapping),  */
/* :136-150 (per-view scatter into the equirect probe, final division).  Per view: img [h,w,3]    */
/* fp32 and bg [h,w] uint8 (non-zero = background pixel) on the device, kinv / rinv = 9 doubles   */
/* on the HOST (inverse intrinsics, inverse pose rotation).  numpy's fancy-index "+=" adds only the */
/* LAST pixel (row-major order) that maps to a texel and counts one hit per view: reproduced with   */
/* an atomicMax bid per texel.  texel [h*w] int32 scratch; winner [lp_h*lp_w] int32 = -1 before the */
/* first view (restored by every call); env [lp_h,lp_w,3] fp64, count [lp_h,lp_w,3] fp32 zeroed by  */
/* the caller, accumulated over views; rnr_stitch_finish divides and writes mask (255 = hit).      */
/* ------------------------------------------------------------------------------------------ */
int rnr_stitch_view(const float* img, const unsigned char* bg, const double* kinv, const double* rinv, int h, int w, int lp_h,
                    int lp_w, int* texel, int* winner, double* env, float* count, void* stream);
int rnr_stitch_finish(double* env, const float* count, unsigned char* mask, int lp_h, int lp_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RNR_B200_H */
