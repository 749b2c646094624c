// Neural-texture sampling: network.TextureMapper.forward (network.py:67-91) on top of
// misc.interpolate_bilinear (misc.py:5-42), forward and backward (texture gradients only -- uv_map
// and sh_basis_map are data, SURVEY.md 8a "Gradient flow"), plus the generic bilinear gather used by
// network.Interpolater (network.py:322-337) and TextureMapper.flatten_mipmap (network.py:93-99).
//
// HBM-bound: per pixel 44 B in (uv + sh basis), 4*C B out; the 4 mip levels (33 MB total for
// C=24) stay L2-resident.  One thread per pixel, float4 texel reads (textures are channels-last).
#include "pixel.cuh"

#define RNR_MAX_LEVELS 8

struct TexLevels {
    const float* tex[RNR_MAX_LEVELS];
    float* gtex[RNR_MAX_LEVELS];
    int size[RNR_MAX_LEVELS];
    int n;
};

namespace {

template <int CMAX>
__global__ void __launch_bounds__(128) texmap_fwd_kernel(const TexLevels lv, int C, const float* __restrict__ uv,
                                                       const float* __restrict__ sh, int sh_start,
                                                       float* __restrict__ out, int64_t HW, int N) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW * N) return;
    const int n = (int)(pix / HW);
    const int64_t p = pix % HW;
    const float u = uv[pix * 2 + 0], v = uv[pix * 2 + 1];
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; c++) acc[c] = 0.f;
    for (int l = 0; l < lv.n; l++) {
        const int S = lv.size[l];
        const float x = u * (float)(S - 1);
        const float y = (float)(S - 1) - v * (float)(S - 1);
        const Bilin b = bilinear_setup(x, y, S, S);
        const float* T = lv.tex[l];
        const float* t00 = T + (int64_t)b.i00 * C;
        const float* t10 = T + (int64_t)b.i10 * C;
        const float* t01 = T + (int64_t)b.i01 * C;
        const float* t11 = T + (int64_t)b.i11 * C;
        if ((C & 3) == 0) {
#pragma unroll
            for (int c = 0; c < CMAX; c += 4) {
                if (c < C) {
                    const float4 a = *(const float4*)(t00 + c), bb = *(const float4*)(t10 + c);
                    const float4 cc = *(const float4*)(t01 + c), d = *(const float4*)(t11 + c);
                    acc[c + 0] += a.x * b.w00 + bb.x * b.w10 + cc.x * b.w01 + d.x * b.w11;
                    acc[c + 1] += a.y * b.w00 + bb.y * b.w10 + cc.y * b.w01 + d.y * b.w11;
                    acc[c + 2] += a.z * b.w00 + bb.z * b.w10 + cc.z * b.w01 + d.z * b.w11;
                    acc[c + 3] += a.w * b.w00 + bb.w * b.w10 + cc.w * b.w01 + d.w * b.w11;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CMAX; c++)
                if (c < C) acc[c] += t00[c] * b.w00 + t10[c] * b.w10 + t01[c] * b.w01 + t11[c] * b.w11;
        }
    }
    if (sh) {
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float s = sh[pix * 9 + k];
#pragma unroll
            for (int c = 0; c < CMAX; c++)
                if (c == sh_start + k) acc[c] *= s;
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; c++)
        if (c < C) out[((int64_t)n * C + c) * HW + p] = acc[c];
}

// Backward of the texture sampling: scatter-add of gout * bilinear weight into the 4 mip levels.
// 262 144 pixels x 4 levels x 4 taps x C channels of atomics funnel into as few as 64^2 texels, so contributions are
// aggregated in the warp first.  A warp owns an 8x4 pixel block (neighbouring pixels hit the same / neighbouring
// texels); for every distinct texel the block touches at a level, each lane sums the weights of its taps on that
// texel, and the C-channel products are reduced across the warp by a transposing butterfly (31 shuffles for 32
// channels, lane c ends up with the sum of channel c) followed by ONE coalesced red.global.add of C floats.
// Texels no pixel maps to receive no atomic at all, so their gradient stays exactly 0 (the albedo-mean loss of
// train_rnr.py:598 relies on bit-identical untouched texels).
template <int CMAX>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[CMAX], int lane) {
    // after the call lane c (c < CMAX) holds sum over lanes of v[c]; CMAX in {16, 32}
    if (CMAX == 32) {
        const bool hi = lane & 16;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float keep = hi ? v[j + 16] : v[j], send = hi ? v[j] : v[j + 16];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    } else {
        // 16 channels: fold the two half-warps first (both halves end with the full sums of 16 channels)
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float keep = hi ? v[j + 8] : v[j], send = hi ? v[j] : v[j + 8];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float keep = hi ? v[j + 4] : v[j], send = hi ? v[j] : v[j + 4];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float keep = hi ? v[j + 2] : v[j], send = hi ? v[j] : v[j + 2];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    {
        const bool hi = lane & 1;
        const float keep = hi ? v[1] : v[0], send = hi ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    return v[0];
}

template <int CMAX>
__global__ void __launch_bounds__(128) texmap_bwd_kernel(const TexLevels lv, int C, const float* __restrict__ uv,
                                                       const float* __restrict__ sh, int sh_start,
                                                       const float* __restrict__ gout, int H, int W, int N) {
    pdl_launch_dependents();
    pdl_wait();
    // block = 4 warps = 16 x 8 pixels; warp = 8 x 4 pixels
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const int n = blockIdx.z;
    const int64_t HW = (int64_t)H * W;
    const bool inb = x < W && y < H;
    const int64_t p = inb ? (int64_t)y * W + x : 0;
    const int64_t pix = (int64_t)n * HW + p;
    float g[CMAX];
    bool any = false;
#pragma unroll
    for (int c = 0; c < CMAX; c++) g[c] = (inb && c < C) ? gout[((int64_t)n * C + c) * HW + p] : 0.f;
    if (sh && inb) {
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float s = sh[pix * 9 + k];
#pragma unroll
            for (int c = 0; c < CMAX; c++)
                if (c == sh_start + k) g[c] *= s;
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; c++) any |= (g[c] != 0.f);       // masked-out pixels contribute exact zeros: no atomics
    if (!__any_sync(0xffffffffu, any)) return;
    float u = 0.f, v = 0.f;
    if (inb) { u = uv[pix * 2 + 0]; v = uv[pix * 2 + 1]; }
    // which of the 32 lanes is responsible for which channel after the transposing reduction
    const int my_c = (CMAX == 32) ? lane : (lane & 15);
    const bool writer = (CMAX == 32) ? (lane < C) : (lane < 16 && lane < C);
    for (int l = 0; l < lv.n; l++) {
        float* T = lv.gtex[l];
        if (!T) continue;
        const int S = lv.size[l];
        const Bilin b = bilinear_setup(u * (float)(S - 1), (float)(S - 1) - v * (float)(S - 1), S, S);
        int key[4] = {b.i00, b.i10, b.i01, b.i11};
        float w[4] = {b.w00, b.w10, b.w01, b.w11};
        if (!any) { w[0] = w[1] = w[2] = w[3] = 0.f; }
#pragma unroll
        for (int t = 0; t < 4; t++) {
            while (true) {
                const unsigned pend = __ballot_sync(0xffffffffu, w[t] != 0.f);
                if (!pend) break;
                const int leader = __ffs(pend) - 1;
                const int k = __shfl_sync(0xffffffffu, key[t], leader);
                float wk = 0.f;
#pragma unroll
                for (int tt = 0; tt < 4; tt++)
                    if (key[tt] == k) { wk += w[tt]; w[tt] = 0.f; }      // (weights of coincident taps add: same as separate atomics)
                float prod[CMAX];
#pragma unroll
                for (int c = 0; c < CMAX; c++) prod[c] = g[c] * wk;
                const float sum = warp_transpose_reduce<CMAX>(prod, lane);
                if (writer) atomicAdd(T + (int64_t)k * C + my_c, sum);
            }
        }
    }
}

// generic gather: data [Nd(1 or N), Hd, Wd, C]; x,y [N, M]; out [N, M, C]
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const float* __restrict__ data, int Nd, int Hd, int Wd, int C,
                                                         const float* __restrict__ xs, const float* __restrict__ ys,
                                                         float* __restrict__ out, int64_t M, int N) {
    const int64_t total = (int64_t)N * M * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t pt = i / C;
        const int n = (int)(pt / M);
        const Bilin b = bilinear_setup(xs[pt], ys[pt], Wd, Hd);
        const float* D = data + (Nd == 1 ? 0 : (int64_t)n * Hd * Wd * C);
        out[i] = D[(int64_t)b.i00 * C + c] * b.w00 + D[(int64_t)b.i10 * C + c] * b.w10 + D[(int64_t)b.i01 * C + c] * b.w01 +
                 D[(int64_t)b.i11 * C + c] * b.w11;
    }
}

__global__ void __launch_bounds__(256) bilinear_bwd_kernel(float* __restrict__ gdata, int Nd, int Hd, int Wd, int C,
                                                         const float* __restrict__ xs, const float* __restrict__ ys,
                                                         const float* __restrict__ gout, int64_t M, int N) {
    const int64_t total = (int64_t)N * M * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t pt = i / C;
        const int n = (int)(pt / M);
        const float g = gout[i];
        if (g == 0.f) continue;
        const Bilin b = bilinear_setup(xs[pt], ys[pt], Wd, Hd);
        float* D = gdata + (Nd == 1 ? 0 : (int64_t)n * Hd * Wd * C);
        if (b.w00 != 0.f) atomicAdd(D + (int64_t)b.i00 * C + c, g * b.w00);
        if (b.w10 != 0.f) atomicAdd(D + (int64_t)b.i10 * C + c, g * b.w10);
        if (b.w01 != 0.f) atomicAdd(D + (int64_t)b.i01 * C + c, g * b.w01);
        if (b.w11 != 0.f) atomicAdd(D + (int64_t)b.i11 * C + c, g * b.w11);
    }
}

}  // namespace

static int fill_levels(TexLevels& lv, const float* const* tex, float* const* gtex, const int* sizes, int L) {
    RNR_REQUIRE(L >= 1 && L <= RNR_MAX_LEVELS, "texture mapper: 1..%d mip levels supported, got %d", RNR_MAX_LEVELS, L);
    memset(&lv, 0, sizeof(lv));
    lv.n = L;
    for (int i = 0; i < L; i++) {
        lv.tex[i] = tex ? tex[i] : nullptr;
        lv.gtex[i] = gtex ? gtex[i] : nullptr;
        lv.size[i] = sizes[i];
    }
    return 0;
}

extern "C" int rnr_texmap_fwd(const float* const* tex, const int* sizes, int L, int C, const float* uv, const float* sh,
                              int sh_start, float* out_nchw, int N, int H, int W, void* stream) {
    TexLevels lv;
    int rc = fill_levels(lv, tex, nullptr, sizes, L);
    if (rc) return rc;
    RNR_REQUIRE(C >= 1 && C <= 32, "texture mapper: 1..32 channels supported, got %d", C);
    RNR_REQUIRE(!sh || (sh_start >= 0 && sh_start + 9 <= C), "texture mapper: SH channels [%d,%d) exceed C=%d", sh_start, sh_start + 9, C);
    const int64_t HW = (int64_t)H * W;
    const int blocks = rnr_cdiv(HW * N, 128);
    if (C <= 16) RNR_PDL_LAUNCH(texmap_fwd_kernel<16>, blocks, 128, 0, stream, lv, C, uv, sh, sh_start, out_nchw, HW, N);
    else RNR_PDL_LAUNCH(texmap_fwd_kernel<32>, blocks, 128, 0, stream, lv, C, uv, sh, sh_start, out_nchw, HW, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_texmap_bwd(float* const* gtex, const int* sizes, int L, int C, const float* uv, const float* sh,
                              int sh_start, const float* gout_nchw, int N, int H, int W, void* stream) {
    TexLevels lv;
    int rc = fill_levels(lv, nullptr, gtex, sizes, L);
    if (rc) return rc;
    RNR_REQUIRE(C >= 1 && C <= 32, "texture mapper: 1..32 channels supported, got %d", C);
    dim3 blocks(rnr_cdiv(W, 16), rnr_cdiv(H, 8), N);
    if (C <= 16) RNR_PDL_LAUNCH(texmap_bwd_kernel<16>, blocks, 128, 0, stream, lv, C, uv, sh, sh_start, gout_nchw, H, W, N);
    else RNR_PDL_LAUNCH(texmap_bwd_kernel<32>, blocks, 128, 0, stream, lv, C, uv, sh, sh_start, gout_nchw, H, W, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bilinear_fwd(const float* data, int Nd, int Hd, int Wd, int C, const float* xs, const float* ys,
                                float* out, int64_t M, int N, void* stream) {
    RNR_REQUIRE(Nd == 1 || Nd == N, "interpolate: data batch must be 1 or N");
    const int64_t total = (int64_t)N * M * C;
    if (total == 0) return 0;
    int blocks = rnr_cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    bilinear_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, Nd, Hd, Wd, C, xs, ys, out, M, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_bilinear_bwd(float* gdata, int Nd, int Hd, int Wd, int C, const float* xs, const float* ys,
                                const float* gout, int64_t M, int N, void* stream) {
    RNR_REQUIRE(Nd == 1 || Nd == N, "interpolate: data batch must be 1 or N");
    const int64_t total = (int64_t)N * M * C;
    if (total == 0) return 0;
    int blocks = rnr_cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    bilinear_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gdata, Nd, Hd, Wd, C, xs, ys, gout, M, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// TextureMapper.flatten_mipmap (network.py:93-99): level 0 + F.interpolate(bilinear, align_corners=False)
// of the coarser levels, channels [c0, c0+nc).  out [1,S0,S0,nc].  Backward scatters into the levels.
// ---------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void up_coord(int dst, int in, int out, int& i0, int& i1, float& l1) {
    const float scale = (float)in / (float)out;
    float src = ((float)dst + 0.5f) * scale - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = src - (float)i0;
}

__global__ void __launch_bounds__(256) flatten_mipmap_kernel(const TexLevels lv, int C, int c0, int nc, float* __restrict__ out,
                                                           const float* __restrict__ gout, int backward) {
    pdl_launch_dependents();
    pdl_wait();
    const int S0 = lv.size[0];
    const int64_t total = (int64_t)S0 * S0 * nc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % nc);
        const int x = (int)((i / nc) % S0), y = (int)(i / ((int64_t)nc * S0));
        float acc = 0.f;
        const float g = backward ? gout[i] : 0.f;
        if (!backward) acc = lv.tex[0][((int64_t)y * S0 + x) * C + c0 + c];
        else if (lv.gtex[0]) atomicAdd(lv.gtex[0] + ((int64_t)y * S0 + x) * C + c0 + c, g);
        for (int l = 1; l < lv.n; l++) {
            const int S = lv.size[l];
            int y0, y1, x0, x1;
            float ly, lx;
            up_coord(y, S, S0, y0, y1, ly);
            up_coord(x, S, S0, x0, x1, lx);
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
            const int64_t o00 = ((int64_t)y0 * S + x0) * C + c0 + c, o01 = ((int64_t)y0 * S + x1) * C + c0 + c;
            const int64_t o10 = ((int64_t)y1 * S + x0) * C + c0 + c, o11 = ((int64_t)y1 * S + x1) * C + c0 + c;
            if (!backward) {
                const float* T = lv.tex[l];
                acc += w00 * T[o00] + w01 * T[o01] + w10 * T[o10] + w11 * T[o11];
            } else if (lv.gtex[l]) {
                float* T = lv.gtex[l];
                atomicAdd(T + o00, g * w00); atomicAdd(T + o01, g * w01);
                atomicAdd(T + o10, g * w10); atomicAdd(T + o11, g * w11);
            }
        }
        if (!backward) out[i] = acc;
    }
}

}  // namespace

extern "C" int rnr_flatten_mipmap(const float* const* tex, float* const* gtex, const int* sizes, int L, int C, int c0, int nc,
                                  float* out, const float* gout, int backward, void* stream) {
    TexLevels lv;
    int rc = fill_levels(lv, tex, gtex, sizes, L);
    if (rc) return rc;
    RNR_REQUIRE(c0 >= 0 && c0 + nc <= C, "flatten_mipmap: channel slice [%d,%d) outside C=%d", c0, c0 + nc, C);
    const int64_t total = (int64_t)sizes[0] * sizes[0] * nc;
    int blocks = rnr_cdiv(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RNR_PDL_LAUNCH(flatten_mipmap_kernel, blocks, 256, 0, stream, lv, C, c0, nc, out, gout, backward);
    RNR_LAUNCH_CHECK();
    return 0;
}
