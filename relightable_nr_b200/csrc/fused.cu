// Fused producers / consumers around the U-Net of the RNR step: the per-pixel stages of train_rnr.py:512-585 written so
// that nothing crosses HBM in a layout the next kernel cannot use directly.
//
//   rnr_head_fwd : TextureMapper.forward (network.py:67-91) + 2x RaySampler.forward (network.py:445-472) + the
//                  torch.cat input assembly (train_rnr.py:530-533) -> the first convolution's operand, i.e. the fp16
//                  channels-last tile with reflect halo that the TMA box loads of conv 'in' read (plus the bf16 copy for
//                  the weight-gradient MMA), the ray uv's and the 6 albedo channels for the tail.  Replaces 3 kernels +
//                  6 permute/cat copies + the NCHW->NHWC pack (436 B/pixel fp32 written and re-read) by one pass.
//   rnr_tail_fwd : rays_lt = (tanh*0.5+0.5)*2 (train_rnr.py:535-536), RayRenderer.forward (network.py:481-527),
//                  RaysLTChromLoss (network.py:395-411) and the cropped masked L1 (train_rnr.py:565-585), read straight
//                  from the last convolution's NHWC output.
//   rnr_tail_bwd : backward of all of the above down to d loss / d (pre-tanh output) in the data-gradient kernel's
//                  operand layout (bf16, zero halo), the last layer's bias gradient, the albedo gradients (planar, for the
//                  texture scatter) and the envmap gradient (one 128-bit vector red per tap).
// The arithmetic of every stage is the one of the module-level kernels (texture.cu, rays.cu, loss.cu), statement by
// statement, so the fused step and the drop-in modules agree to fp32 rounding.
#include "pixel.cuh"
#include <stdlib.h>

namespace {

constexpr int kHeadPx = 32;
constexpr int kHeadThreads = 256;
constexpr int kMaxRays = 32;

struct HeadParams {
    const float* tex[8];
    int size[8];
    int n_levels, C, sh_start;
    const float* uv;         // [P,2]
    const float* sh;         // [P,9] or null
    const float* tbn;        // [P,9]
    const float* vdt;        // [P,3]
    const float* alpha;      // [P]
    const float* normal;     // [P,3] or null
    const float* view_dir;   // [P,3] or null
    const float* piv_s;      // [3,Rs]
    const float* piv_d;      // [3,Rd]
    int Rs, Rd;
    __half* act;             // [N,H+2,W+2,Cpad]
    __nv_bfloat16* act_b;    // same or null
    int Cpad;
    float* rays_uv;          // [P,2,R] or null
    float* albedo;           // [P,8] or null
    int N, H, W;
};

__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(const HeadParams q) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    const int R = q.Rs + q.Rd;
    float* s_out = sm;                                   // [kHeadPx][Cpad]
    float* s_uv = s_out + kHeadPx * q.Cpad;              // [kHeadPx][2R]
    float* s_in = s_uv + kHeadPx * 2 * kMaxRays;         // [kHeadPx][13]: TBN 9, vdt 3, alpha
    float* s_piv = s_in + kHeadPx * 13;                  // [2][3*kMaxRays]
    const int tid = threadIdx.x;
    const int64_t P = (int64_t)q.N * q.H * q.W;
    const int64_t p0 = (int64_t)blockIdx.x * kHeadPx;
    const int np = (int)((P - p0) < kHeadPx ? (P - p0) : kHeadPx);

    for (int i = tid; i < kHeadPx * q.Cpad; i += kHeadThreads) s_out[i] = 0.f;
    for (int i = tid; i < 3 * q.Rs; i += kHeadThreads) s_piv[i] = q.piv_s[i];
    for (int i = tid; i < 3 * q.Rd; i += kHeadThreads) s_piv[3 * kMaxRays + i] = q.piv_d[i];
    if (R > 0) {
        for (int i = tid; i < np * 9; i += kHeadThreads) s_in[(i / 9) * 13 + (i % 9)] = q.tbn[p0 * 9 + i];
        for (int i = tid; i < np * 3; i += kHeadThreads) s_in[(i / 3) * 13 + 9 + (i % 3)] = q.vdt[p0 * 3 + i];
        for (int i = tid; i < np; i += kHeadThreads) s_in[i * 13 + 12] = q.alpha[p0 + i];
    }
    __syncthreads();

    // ---- rays: item = (pixel, ray) ----
    for (int i = tid; i < np * R; i += kHeadThreads) {
        const int px = i / R, r = i - px * R;
        const float* T = s_in + px * 13;
        const float a = T[12];
        const bool reflect = r < q.Rs;
        // background pixel (alpha == 0): the reflected direction is scaled by alpha, so dir = 0 and uv = 0.5 * 0 - 1 exactly; the
        // diffuse pivots give the same when TBN is zero (how the rasterizer / precompute.py leave uncovered pixels).  Skips the
        // two normalisations and atan2 / acos for typically half of the image without changing a bit of the result.
        if (a == 0.f && (reflect || (T[0] == 0.f && T[1] == 0.f && T[2] == 0.f && T[3] == 0.f && T[4] == 0.f && T[5] == 0.f &&
                                     T[6] == 0.f && T[7] == 0.f && T[8] == 0.f))) {
            s_uv[px * 2 * R + r] = -1.f;
            s_uv[px * 2 * R + R + r] = -1.f;
            continue;                                   // (s_out is pre-zeroed)
        }
        const float* pv = reflect ? s_piv : s_piv + 3 * kMaxRays;
        const int Rm = reflect ? q.Rs : q.Rd, rr = reflect ? r : r - q.Rs;
        const float px_ = pv[rr], py_ = pv[Rm + rr], pz_ = pv[2 * Rm + rr];
        float tx, ty, tz;
        if (reflect) {
            const float vx = T[9], vy = T[10], vz = T[11];
            const float d = (px_ * vx + py_ * vy + pz_ * vz) * 2.0f;
            tx = d * px_ - vx; ty = d * py_ - vy; tz = d * pz_ - vz;
            normalize3(tx, ty, tz);
            tx *= a; ty *= a; tz *= a;
        } else {
            tx = px_; ty = py_; tz = pz_;
        }
        float wx = T[0] * tx + T[1] * ty + T[2] * tz;
        float wy = T[3] * tx + T[4] * ty + T[5] * tz;
        float wz = T[6] * tx + T[7] * ty + T[8] * tz;
        normalize3(wx, wy, wz);
        float* o = s_out + px * q.Cpad + r * 3;
        o[0] = wx; o[1] = wy; o[2] = wz;
        float u, v;
        spherical_uv(wx, wy, wz, u, v);
        const float bg = (a == 0.f) ? 1.f : 0.f;
        s_uv[px * 2 * R + r] = u * a - bg;
        s_uv[px * 2 * R + R + r] = v * a - bg;
    }
    // ---- normal / view direction ----
    int cbase = 3 * R;
    if (q.normal) {
        for (int i = tid; i < np * 3; i += kHeadThreads) s_out[(i / 3) * q.Cpad + cbase + (i % 3)] = q.normal[p0 * 3 + i];
        cbase += 3;
    }
    if (q.view_dir) {
        for (int i = tid; i < np * 3; i += kHeadThreads) s_out[(i / 3) * q.Cpad + cbase + (i % 3)] = q.view_dir[p0 * 3 + i];
        cbase += 3;
    }
    // ---- neural texture: item = (pixel, group of 4 channels) ----
    const int ng = q.C >> 2;
    for (int i = tid; i < np * ng; i += kHeadThreads) {
        const int px = i / ng, c = (i - px * ng) * 4;
        const int64_t pix = p0 + px;
        const float u = q.uv[pix * 2 + 0], v = q.uv[pix * 2 + 1];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < q.n_levels; l++) {
            const int S = q.size[l];
            const float x = u * (float)(S - 1);
            const float y = (float)(S - 1) - v * (float)(S - 1);
            const Bilin b = bilinear_setup(x, y, S, S);
            const float* Tx = q.tex[l];
            const float4 a = *(const float4*)(Tx + (int64_t)b.i00 * q.C + c), bb = *(const float4*)(Tx + (int64_t)b.i10 * q.C + c);
            const float4 cc = *(const float4*)(Tx + (int64_t)b.i01 * q.C + c), d = *(const float4*)(Tx + (int64_t)b.i11 * q.C + c);
            acc[0] += a.x * b.w00 + bb.x * b.w10 + cc.x * b.w01 + d.x * b.w11;
            acc[1] += a.y * b.w00 + bb.y * b.w10 + cc.y * b.w01 + d.y * b.w11;
            acc[2] += a.z * b.w00 + bb.z * b.w10 + cc.z * b.w01 + d.z * b.w11;
            acc[3] += a.w * b.w00 + bb.w * b.w10 + cc.w * b.w01 + d.w * b.w11;
        }
        if (q.sh) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = c + e - q.sh_start;
                if (k >= 0 && k < 9) acc[e] *= q.sh[pix * 9 + k];
            }
        }
        float* o = s_out + px * q.Cpad + cbase + c;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
        if (q.albedo && c < 8) *(float4*)(q.albedo + pix * 8 + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    __syncthreads();

    // ---- write-out: fp16 (+ bf16) rows of Cpad channels with the reflect halo; rays_uv rows ----
    const int vpp = q.Cpad >> 3;
    const int Hp = q.H + 2, Wp = q.W + 2;
    for (int i = tid; i < np * vpp; i += kHeadThreads) {
        const int px = i / vpp, c = (i - px * vpp) * 8;
        const int64_t pix = p0 + px;
        const int w = (int)(pix % q.W), h = (int)((pix / q.W) % q.H), n = (int)(pix / ((int64_t)q.W * q.H));
        const float* sv = s_out + px * q.Cpad + c;
        __align__(16) __half o[8];
        __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { o[e] = __float2half_rn(sv[e]); ob[e] = __float2bfloat16_rn(sv[e]); }
        const uint4 ov = *(const uint4*)o, ovb = *(const uint4*)ob;
        int rows[3], cols[3], nr = 0, nc = 0;
        rows[nr++] = h + 1;
        if (h == 1) rows[nr++] = 0;
        if (h == q.H - 2) rows[nr++] = q.H + 1;
        cols[nc++] = w + 1;
        if (w == 1) cols[nc++] = 0;
        if (w == q.W - 2) cols[nc++] = q.W + 1;
        for (int a = 0; a < nr; a++)
            for (int b = 0; b < nc; b++) {
                const int64_t o_ = (((int64_t)n * Hp + rows[a]) * Wp + cols[b]) * q.Cpad + c;
                *(uint4*)(q.act + o_) = ov;
                if (q.act_b) *(uint4*)(q.act_b + o_) = ovb;
            }
    }
    if (q.rays_uv)
        for (int i = tid; i < np * 2 * R; i += kHeadThreads) q.rays_uv[p0 * 2 * R + i] = s_uv[i];
}

// -------------------------------------------------------------------------------------------------------------------
// tail
// -------------------------------------------------------------------------------------------------------------------
struct TailParams {
    const float* raw;        // [P, ldraw] tanh output of the last convolution (channels r*3+c)
    int ldraw;
    const float* rays_uv;    // [P,2,R]
    const float* albedo;     // [P,8]: diffuse 0..2, specular 3..5
    const float4* lp4;       // [Hl*Wl] envmap texels (r, g, b, unused): one 16-byte load per bilinear tap
    int Hl, Wl;
    const float* alpha;      // [P]
    const float* img;        // [N,3,H,W]
    int R, Rs, Rd;
    int N, H, W, crop;
    float* final_img;        // [N,3,H,W]
    float* aux;              // [P,12]: ltt_s 0..2, ltt_d 3..5, chrom mean 6..8, image weight 9
    double* sums;            // [0] sum diff, [1] sum alpha, [2] sum |out*a - gt*a| over the crop
};

constexpr int kTailPx = 32;             // pixels per group
constexpr int kTailThreads = 256;       // work item = (pixel, ray): 32 x 26 items per group, ~3 per thread
// row pitches in shared memory: smallest odd numbers > 3R resp. 2R
__host__ __device__ __forceinline__ int raw_pitch(int R) { return (3 * R + 1) | 1; }
__host__ __device__ __forceinline__ int uv_pitch(int R) { return (2 * R + 1) | 1; }

// Coalesced staging of a group's rows of the last convolution's output ([P, ldraw], 3R used) and of rays_uv ([P, 2R]) into
// shared memory.  (Reading the rows straight from global memory, one thread per 312-byte row, moved 6x the useful bytes
// through L1; one thread per PIXEL left 12 warps per SM walking 26 dependent rays each: the kernels ran at 30 % issue
// utilisation.  Items are now (pixel, ray) pairs, 2048 threads per SM.)
__device__ __forceinline__ void stage_rows(const TailParams& q, int64_t p0, int np, float* s_raw, float* s_uv) {
    const int kRawPitch = raw_pitch(q.R), kUvPitch = uv_pitch(q.R);
    const int nch = 3 * q.R, nv = (nch + 3) >> 2;
    for (int i = threadIdx.x; i < np * nv; i += kTailThreads) {
        const int px = i / nv, k = (i - px * nv) * 4;
        const float4 v = *(const float4*)(q.raw + (p0 + px) * q.ldraw + k);
        float* d = s_raw + px * kRawPitch + k;
        d[0] = v.x; if (k + 1 < kRawPitch) d[1] = v.y; if (k + 2 < kRawPitch) d[2] = v.z; if (k + 3 < kRawPitch) d[3] = v.w;
    }
    const int nuv = 2 * q.R;
    for (int i = threadIdx.x; i < np * nuv; i += kTailThreads) {
        const int px = i / nuv, k = i - px * nuv;
        s_uv[px * kUvPitch + k] = q.rays_uv[p0 * nuv + i];
    }
}

__device__ __forceinline__ void env_taps(const TailParams& q, const float* uvrow, int r, Bilin& b) {
    const float u = uvrow[r], v = uvrow[q.R + r];
    const float x = fminf(u * (float)q.Wl, (float)(q.Wl - 1));
    const float y = fminf(v * (float)q.Hl, (float)(q.Hl - 1));
    b = bilinear_setup(x, y, q.Wl, q.Hl);
}

__device__ __forceinline__ void env_color(const TailParams& q, const Bilin& b, float* col) {
    const float4 t00 = q.lp4[b.i00], t10 = q.lp4[b.i10], t01 = q.lp4[b.i01], t11 = q.lp4[b.i11];
    col[0] = t00.x * b.w00 + t10.x * b.w10 + t01.x * b.w01 + t11.x * b.w11;
    col[1] = t00.y * b.w00 + t10.y * b.w10 + t01.y * b.w01 + t11.y * b.w11;
    col[2] = t00.z * b.w00 + t10.z * b.w10 + t01.z * b.w01 + t11.z * b.w11;
}

__device__ __forceinline__ float block_sum256(float v, float* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_tmp[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = lane < (kTailThreads >> 5) ? s_tmp[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;   // valid in warp 0
}

__global__ void __launch_bounds__(kTailThreads) tail_fwd_kernel(const TailParams q) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_dyn[];
    const int kRawPitch = raw_pitch(q.R), kUvPitch = uv_pitch(q.R);
    float* s_raw = s_dyn;                              // [32][kRawPitch]: tanh outputs, then rays_lt * envmap colour
    float* s_uv = s_raw + kTailPx * kRawPitch;         // [32][kUvPitch]
    float* s_ch = s_uv + kTailPx * kUvPitch;           // [32][kRawPitch]: normalised rays_lt (chromaticity)
    float* s_m = s_ch + kTailPx * kRawPitch;           // [32][4]: normalised mean chromaticity, alpha * image weight
    __shared__ float s_tmp[8];
    const int64_t HW = (int64_t)q.H * q.W;
    const int64_t P = HW * q.N;
    const int ngroups = (int)((P + kTailPx - 1) / kTailPx);
    const int tid = threadIdx.x;
    float dsum = 0.f, asum = 0.f, lsum = 0.f;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t p0 = (int64_t)grp * kTailPx;
        const int np = (int)((P - p0) < kTailPx ? (P - p0) : kTailPx);
        __syncthreads();                               // previous group's readers are done with the shared buffers
        stage_rows(q, p0, np, s_raw, s_uv);
        __syncthreads();
        // ---- (pixel, ray): envmap colour, rays_lt, chromaticity ----
        for (int i = tid; i < np * q.R; i += kTailThreads) {
            const int px = i / q.R, r = i - px * q.R;
            Bilin b;
            env_taps(q, s_uv + px * kUvPitch, r, b);
            float col[3], lt[3];
            env_color(q, b, col);
            float* t = s_raw + px * kRawPitch + r * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) { lt[c] = (t[c] * 0.5f + 0.5f) * 2.0f; t[c] = lt[c] * col[c]; }
            float x = lt[0], y = lt[1], z = lt[2];
            normalize3(x, y, z);
            float* ch = s_ch + px * kRawPitch + r * 3;
            ch[0] = x; ch[1] = y; ch[2] = z;
        }
        __syncthreads();
        // ---- per pixel: ray sums (same summation order as the module kernels), image, L1, aux ----
        if (tid < np) {
            const int px = tid;
            const int64_t pix = p0 + px;
            const int n = (int)(pix / HW);
            const int64_t p = pix % HW;
            const float a = q.alpha[pix];
            float gt[3];
#pragma unroll
            for (int c = 0; c < 3; c++) gt[c] = q.img[((int64_t)n * 3 + c) * HW + p];
            const float wi = fminf(sqrtf(gt[0] * gt[0] + gt[1] * gt[1] + gt[2] * gt[2]) * 20.f, 1.0f);
            const float* tt = s_raw + px * kRawPitch;
            const float* ch = s_ch + px * kRawPitch;
            float ss[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
            float m0 = 0.f, m1 = 0.f, m2 = 0.f;
            for (int r = 0; r < q.R; r++) {
#pragma unroll
                for (int c = 0; c < 3; c++) { if (r < q.Rs) ss[c] += tt[r * 3 + c]; else sd[c] += tt[r * 3 + c]; }
                m0 += ch[r * 3 + 0]; m1 += ch[r * 3 + 1]; m2 += ch[r * 3 + 2];
            }
            m0 /= (float)q.R; m1 /= (float)q.R; m2 /= (float)q.R;
            normalize3(m0, m1, m2);
            s_m[px * 4 + 0] = m0; s_m[px * 4 + 1] = m1; s_m[px * 4 + 2] = m2; s_m[px * 4 + 3] = a * wi;
            asum += a;
            const int w = (int)(p % q.W), h = (int)(p / q.W);
            const bool inside = h >= q.crop && h < q.H - q.crop && w >= q.crop && w < q.W - q.crop;
            const float4 al0 = *(const float4*)(q.albedo + pix * 8), al1 = *(const float4*)(q.albedo + pix * 8 + 4);
            const float adv[3] = {al0.x, al0.y, al0.z}, asv[3] = {al0.w, al1.x, al1.y};
            float axv[12];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float ls = ss[c] / (float)q.Rs;
                const float os = asv[c] * ls;
                float ld = 0.f, od = 0.f;
                if (q.Rd > 0) { ld = sd[c] / (float)q.Rd; od = adv[c] * ld; }
                const float out = os + od;
                q.final_img[((int64_t)n * 3 + c) * HW + p] = out;
                axv[c] = ls; axv[3 + c] = ld;
                if (inside) lsum += fabsf(out * a - gt[c] * a);
            }
            axv[6] = m0; axv[7] = m1; axv[8] = m2; axv[9] = wi; axv[10] = 0.f; axv[11] = 0.f;
            float* ax = q.aux + pix * 12;
            *(float4*)(ax + 0) = make_float4(axv[0], axv[1], axv[2], axv[3]);
            *(float4*)(ax + 4) = make_float4(axv[4], axv[5], axv[6], axv[7]);
            *(float4*)(ax + 8) = make_float4(axv[8], axv[9], axv[10], axv[11]);
        }
        __syncthreads();
        // ---- (pixel, ray): chromaticity deviation from the pixel's mean ----
        for (int i = tid; i < np * q.R; i += kTailThreads) {
            const int px = i / q.R, r = i - px * q.R;
            const float* ch = s_ch + px * kRawPitch + r * 3;
            const float* m = s_m + px * 4;
            dsum += (1.f - (ch[0] * m[0] + ch[1] * m[1] + ch[2] * m[2])) * m[3];
        }
    }
    const float bd = block_sum256(dsum, s_tmp);
    const float ba = block_sum256(asum, s_tmp);
    const float bl = block_sum256(lsum, s_tmp);
    if (threadIdx.x == 0) {
        if (bd != 0.f) atomicAdd(&q.sums[0], (double)bd);
        if (ba != 0.f) atomicAdd(&q.sums[1], (double)ba);
        if (bl != 0.f) atomicAdd(&q.sums[2], (double)bl);
    }
}

struct TailBwdParams {
    TailParams f;
    float w_l1, w_chrom;     // loss weights (upstream gradient folded in)
    __nv_bfloat16* gz;       // [N,H+2,W+2,ldg] zero halo; channels [0, roundup8(3R)) of the interior are written
    int ldg;
    float* dbias;            // [3R] += sum over pixels of dz
    float* g_alb;            // [N,6,H,W]: d loss / d neural_img channels 0..5
    float4* g_lp4;           // [Hl*Wl] (r,g,b,unused) += envmap gradient, or null
};

__global__ void __launch_bounds__(kTailThreads) tail_bwd_kernel(const TailBwdParams qq) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_dyn[];
    const TailParams& q = qq.f;
    const int kRawPitch = raw_pitch(q.R), kUvPitch = uv_pitch(q.R);
    float* s_raw = s_dyn;                              // [32][kRawPitch]: tanh outputs, overwritten in place by dz
    float* s_uv = s_raw + kTailPx * kRawPitch;         // [32][kUvPitch]
    float* s_px = s_uv + kTailPx * kUvPitch;           // [32][12]: Gls 0..2, Gld 3..5, mean chromaticity 6..8, chrom scale 9
    const int64_t HW = (int64_t)q.H * q.W;
    const int64_t P = HW * q.N;
    const int ngroups = (int)((P + kTailPx - 1) / kTailPx);
    const int tid = threadIdx.x;
    const int nch = 3 * q.R;
    const double cnt = (double)q.N * 3 * (q.H - 2 * q.crop) * (q.W - 2 * q.crop);
    const float gl1 = (float)((double)qq.w_l1 / cnt);
    float bias_acc = 0.f;                              // thread c < 3R owns the bias gradient of channel c across its groups
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t p0 = (int64_t)grp * kTailPx;
        const int np = (int)((P - p0) < kTailPx ? (P - p0) : kTailPx);
        __syncthreads();
        stage_rows(q, p0, np, s_raw, s_uv);
        // ---- per pixel: image-loss gradient, albedo gradients, per-ray scale factors ----
        if (tid < np) {
            const int px = tid;
            const int64_t pix = p0 + px;
            const int n = (int)(pix / HW);
            const int64_t p = pix % HW;
            const float a = q.alpha[pix];
            const float4 ax0 = *(const float4*)(q.aux + pix * 12), ax1 = *(const float4*)(q.aux + pix * 12 + 4), ax2 = *(const float4*)(q.aux + pix * 12 + 8);
            const float lsv[3] = {ax0.x, ax0.y, ax0.z}, ldv[3] = {ax0.w, ax1.x, ax1.y};
            const float4 al0 = *(const float4*)(q.albedo + pix * 8), al1 = *(const float4*)(q.albedo + pix * 8 + 4);
            const float adv[3] = {al0.x, al0.y, al0.z}, asv[3] = {al0.w, al1.x, al1.y};
            const int w = (int)(p % q.W), h = (int)(p / q.W);
            const bool inside = h >= q.crop && h < q.H - q.crop && w >= q.crop && w < q.W - q.crop;
            float* sp = s_px + px * 12;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float as = asv[c], ad = adv[c];
                const float ls = lsv[c], ld = ldv[c];
                const float os = as * ls;
                const float od = (q.Rd > 0) ? ad * ld : 0.f;
                const float out = os + od;
                float go = 0.f;
                if (inside) {
                    const float d = out * a - q.img[((int64_t)n * 3 + c) * HW + p] * a;
                    go = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * a * gl1;
                }
                sp[c] = go * as / (float)q.Rs;
                sp[3 + c] = (q.Rd > 0) ? go * ad / (float)q.Rd : 0.f;
                qq.g_alb[((int64_t)n * 6 + 3 + c) * HW + p] = go * ls;
                qq.g_alb[((int64_t)n * 6 + c) * HW + p] = (q.Rd > 0) ? go * ld : 0.f;
            }
            sp[6] = ax1.z; sp[7] = ax1.w; sp[8] = ax2.x;
            sp[9] = qq.w_chrom * a * ax2.y / ((float)q.sums[1] * (float)q.R);
        }
        __syncthreads();
        // ---- (pixel, ray): d loss / d rays_lt -> d loss / d (pre-tanh output), envmap gradient ----
        for (int i = tid; i < np * q.R; i += kTailThreads) {
            const int px = i / q.R, r = i - px * q.R;
            const float* sp = s_px + px * 12;
            Bilin b;
            env_taps(q, s_uv + px * kUvPitch, r, b);
            float col[3];
            env_color(q, b, col);
            float* t = s_raw + px * kRawPitch + r * 3;
            float th[3], lt[3];
#pragma unroll
            for (int c = 0; c < 3; c++) { th[c] = t[c]; lt[c] = (th[c] * 0.5f + 0.5f) * 2.0f; }
            const float nr = fmaxf(sqrtf(lt[0] * lt[0] + lt[1] * lt[1] + lt[2] * lt[2]), 1e-12f);
            const float x = lt[0] / nr, y = lt[1] / nr, z = lt[2] / nr;
            const float m0 = sp[6], m1 = sp[7], m2 = sp[8], sc = sp[9];
            const float cm = x * m0 + y * m1 + z * m2;
            const float gch[3] = {-sc * (m0 - x * cm) / nr, -sc * (m1 - y * cm) / nr, -sc * (m2 - z * cm) / nr};
            float gc[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float G = (r < q.Rs) ? sp[c] : sp[3 + c];
                const float glt = G * col[c] + gch[c];
                t[c] = glt * (1.f - th[c] * th[c]);
                gc[c] = G * lt[c];
            }
            if (qq.g_lp4 && (gc[0] != 0.f || gc[1] != 0.f || gc[2] != 0.f)) {
                if (b.w00 != 0.f) atomicAdd(qq.g_lp4 + b.i00, make_float4(gc[0] * b.w00, gc[1] * b.w00, gc[2] * b.w00, 0.f));
                if (b.w10 != 0.f) atomicAdd(qq.g_lp4 + b.i10, make_float4(gc[0] * b.w10, gc[1] * b.w10, gc[2] * b.w10, 0.f));
                if (b.w01 != 0.f) atomicAdd(qq.g_lp4 + b.i01, make_float4(gc[0] * b.w01, gc[1] * b.w01, gc[2] * b.w01, 0.f));
                if (b.w11 != 0.f) atomicAdd(qq.g_lp4 + b.i11, make_float4(gc[0] * b.w11, gc[1] * b.w11, gc[2] * b.w11, 0.f));
            }
        }
        __syncthreads();
        // ---- gz rows: bf16, 8 channels per 16-byte store, channels [0, roundup8(3R)) ----
        const int vpp = (nch + 7) >> 3;
        const int Hp = q.H + 2, Wp = q.W + 2;
        for (int i = tid; i < np * vpp; i += kTailThreads) {
            const int px = i / vpp, c = (i - px * vpp) * 8;
            const int64_t pp = p0 + px;
            const int w = (int)(pp % q.W), h = (int)((pp / q.W) % q.H), n = (int)(pp / HW);
            __align__(16) __nv_bfloat16 o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = __float2bfloat16_rn(c + e < nch ? s_raw[px * kRawPitch + c + e] : 0.f);
            *(uint4*)(qq.gz + (((int64_t)n * Hp + h + 1) * Wp + w + 1) * qq.ldg + c) = *(const uint4*)o;
        }
        // ---- bias gradient of the last convolution: column sums, kept in a register across the block's groups ----
        if (tid < nch)
            for (int px = 0; px < np; px++) bias_acc += s_raw[px * kRawPitch + tid];
    }
    if (tid < nch && bias_acc != 0.f) atomicAdd(qq.dbias + tid, bias_acc);
}

}  // namespace

extern "C" int rnr_head_fwd(const float* const* tex, const int* sizes, int n_levels, int C, const float* uv, const float* sh,
                            int sh_start, const float* tbn, const float* vdt, const float* alpha, const float* normal,
                            const float* view_dir, const float* pivots_s, int Rs, const float* pivots_d, int Rd, void* act,
                            void* act_bf16, int Cpad, float* rays_uv, float* albedo, int N, int H, int W, void* stream) {
    RNR_REQUIRE(n_levels >= 1 && n_levels <= 8, "head: 1..8 mip levels supported, got %d", n_levels);
    RNR_REQUIRE(C >= 4 && C % 4 == 0 && C <= 64, "head: texture channels must be a multiple of 4 (got %d)", C);
    RNR_REQUIRE(Rs >= 0 && Rd >= 0 && Rs + Rd <= kMaxRays, "head: at most %d rays (got %d + %d)", kMaxRays, Rs, Rd);
    RNR_REQUIRE(!sh || (sh_start >= 0 && sh_start + 9 <= C), "head: SH channels [%d,%d) exceed C=%d", sh_start, sh_start + 9, C);
    const int used = 3 * (Rs + Rd) + (normal ? 3 : 0) + (view_dir ? 3 : 0) + C;
    RNR_REQUIRE(Cpad % 8 == 0 && used <= Cpad && Cpad <= 256, "head: %d channels do not fit the operand pitch %d", used, Cpad);
    RNR_REQUIRE(H >= 2 && W >= 2, "head: reflect halo needs H, W >= 2");
    RNR_REQUIRE(Rs + Rd == 0 || (tbn && vdt && alpha && pivots_s && (Rd == 0 || pivots_d)), "head: ray inputs missing");
    HeadParams q;
    memset(&q, 0, sizeof(q));
    for (int i = 0; i < n_levels; i++) { q.tex[i] = tex[i]; q.size[i] = sizes[i]; }
    q.n_levels = n_levels; q.C = C; q.sh_start = sh_start;
    q.uv = uv; q.sh = sh; q.tbn = tbn; q.vdt = vdt; q.alpha = alpha; q.normal = normal; q.view_dir = view_dir;
    q.piv_s = pivots_s; q.piv_d = pivots_d; q.Rs = Rs; q.Rd = Rd;
    q.act = (__half*)act; q.act_b = (__nv_bfloat16*)act_bf16; q.Cpad = Cpad;
    q.rays_uv = rays_uv; q.albedo = albedo; q.N = N; q.H = H; q.W = W;
    const size_t smem = ((size_t)kHeadPx * Cpad + kHeadPx * 2 * kMaxRays + kHeadPx * 13 + 6 * kMaxRays) * sizeof(float);
    RNR_ONCE_PER_DEVICE({ RNR_CHECK(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); });
    RNR_REQUIRE(smem <= 64 * 1024, "head: shared memory %zu too large", smem);
    const int64_t P = (int64_t)N * H * W;
    RNR_PDL_LAUNCH(head_fwd_kernel, rnr_cdiv(P, kHeadPx), kHeadThreads, smem, stream, q);
    RNR_LAUNCH_CHECK();
    return 0;
}

static int tail_carveout() {
    const char* e = getenv("RNR_TAIL_CARVEOUT");            // percent of the SM's shared memory / L1 array given to shared memory
    const int v = e ? atoi(e) : 50;
    return v < 0 ? 0 : (v > 100 ? 100 : v);
}

static int fill_tail(TailParams& q, const float* raw, int ldraw, const float* rays_uv, const float* albedo, const float* lp4,
                     int Hl, int Wl, const float* alpha, const float* img, int Rs, int Rd, int N, int H, int W, int crop,
                     float* final_img, float* aux, double* sums) {
    RNR_REQUIRE(Rs >= 1 && Rd >= 0 && Rs + Rd <= kMaxRays, "tail: 1..%d rays supported (got %d + %d)", kMaxRays, Rs, Rd);
    RNR_REQUIRE(ldraw >= (3 * (Rs + Rd) + 3) / 4 * 4, "tail: output pitch %d < %d channels", ldraw, 3 * (Rs + Rd));
    RNR_REQUIRE(H > 2 * crop && W > 2 * crop, "tail: crop too large");
    RNR_REQUIRE(raw && rays_uv && albedo && lp4 && alpha && img && aux && sums, "tail: null argument");
    RNR_REQUIRE(ldraw % 4 == 0 && (((uintptr_t)raw | (uintptr_t)lp4 | (uintptr_t)aux | (uintptr_t)albedo) & 15) == 0, "tail: 16-byte alignment required");
    q.raw = raw; q.ldraw = ldraw; q.rays_uv = rays_uv; q.albedo = albedo; q.lp4 = (const float4*)lp4; q.Hl = Hl; q.Wl = Wl; q.alpha = alpha;
    q.img = img; q.R = Rs + Rd; q.Rs = Rs; q.Rd = Rd; q.N = N; q.H = H; q.W = W; q.crop = crop; q.final_img = final_img;
    q.aux = aux; q.sums = sums;
    return 0;
}

extern "C" int rnr_tail_fwd(const float* raw, int ldraw, const float* rays_uv, const float* albedo, const float* lp4, int Hl, int Wl,
                            const float* alpha, const float* img, int Rs, int Rd, int N, int H, int W, int crop, float* final_img,
                            float* aux, double* sums, void* stream) {
    TailParams q;
    int rc = fill_tail(q, raw, ldraw, rays_uv, albedo, lp4, Hl, Wl, alpha, img, Rs, Rd, N, H, W, crop, final_img, aux, sums);
    if (rc) return rc;
    RNR_REQUIRE(final_img, "tail: null output");
    const size_t smem = (size_t)kTailPx * (2 * raw_pitch(Rs + Rd) + uv_pitch(Rs + Rd) + 4) * sizeof(float);
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        // the kernel is bound by the envmap texel gathers (4 x 16 B per ray, L1 misses go to L2 one sector at a time): cap the
        // shared-memory carve-out so that the rest of the 228 KB stays L1 for the 2 MB envmap's working set
        RNR_CHECK(cudaFuncSetAttribute(tail_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, tail_carveout()));
    });
    int blocks = rnr_cdiv((int64_t)N * H * W, kTailPx);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RNR_PDL_LAUNCH(tail_fwd_kernel, blocks, kTailThreads, smem, stream, q);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_tail_bwd(const float* raw, int ldraw, const float* rays_uv, const float* albedo, const float* lp4, int Hl, int Wl,
                            const float* alpha, const float* img, int Rs, int Rd, int N, int H, int W, int crop, const float* aux,
                            const double* sums, float w_l1, float w_chrom, void* gz, int ldg, float* dbias, float* g_alb,
                            float* g_lp4, void* stream) {
    TailBwdParams qq;
    int rc = fill_tail(qq.f, raw, ldraw, rays_uv, albedo, lp4, Hl, Wl, alpha, img, Rs, Rd, N, H, W, crop, nullptr,
                       const_cast<float*>(aux), const_cast<double*>(sums));
    if (rc) return rc;
    RNR_REQUIRE(gz && dbias && g_alb, "tail bwd: null output");
    RNR_REQUIRE(ldg % 8 == 0 && ldg >= (3 * (Rs + Rd) + 7) / 8 * 8, "tail bwd: gradient pitch %d too small", ldg);
    RNR_REQUIRE(!g_lp4 || ((uintptr_t)g_lp4 & 15) == 0, "tail bwd: envmap gradient must be 16-byte aligned");
    qq.w_l1 = w_l1; qq.w_chrom = w_chrom; qq.gz = (__nv_bfloat16*)gz; qq.ldg = ldg; qq.dbias = dbias; qq.g_alb = g_alb;
    qq.g_lp4 = (float4*)g_lp4;
    const size_t smem = (size_t)kTailPx * (raw_pitch(Rs + Rd) + uv_pitch(Rs + Rd) + 12) * sizeof(float);
    RNR_ONCE_PER_DEVICE({
        RNR_CHECK(cudaFuncSetAttribute(tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        RNR_CHECK(cudaFuncSetAttribute(tail_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, tail_carveout()));
    });
    int blocks = rnr_cdiv((int64_t)N * H * W, kTailPx);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RNR_PDL_LAUNCH(tail_bwd_kernel, blocks, kTailThreads, smem, stream, qq);
    RNR_LAUNCH_CHECK();
    return 0;
}
