// Training-step tail: losses and the optimiser.
//   network.RaysLTChromLoss.forward          network.py:395-411 (forward + backward w.r.t. rays_lt)
//   masked / cropped L1 image loss            train_rnr.py:565-585 (criterionL1 on outputs*alpha vs gt*alpha, 5 px crop)
//   torch.optim.Adam(lr, betas, eps) step     train_rnr.py:376,622  (fused over one flat parameter buffer)
#include "pixel.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v, float* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_tmp[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? s_tmp[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;   // valid in warp 0
}

// ---- RaysLTChromLoss ------------------------------------------------------------------------
// sums[0] += sum of diff, sums[1] += sum of alpha.  Optional full outputs for the module API.
__global__ void __launch_bounds__(128) chrom_fwd_kernel(const float* __restrict__ lt, const float* __restrict__ alpha,
                                                      const float* __restrict__ img, int R, int64_t HW, int N,
                                                      float* __restrict__ chrom, float* __restrict__ chrom_mean,
                                                      float* __restrict__ diff, double* __restrict__ sums) {
    __shared__ float s_tmp[8];
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float dsum = 0.f, asum = 0.f;
    if (pix < HW * N) {
        const int n = (int)(pix / HW);
        const int64_t p = pix % HW;
        const float a = alpha[pix];
        float w = 1.f;
        if (img) {
            const float i0 = img[((int64_t)n * 3 + 0) * HW + p], i1 = img[((int64_t)n * 3 + 1) * HW + p], i2 = img[((int64_t)n * 3 + 2) * HW + p];
            w = fminf(sqrtf(i0 * i0 + i1 * i1 + i2 * i2) * 20.f, 1.0f);
        }
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        for (int r = 0; r < R; r++) {
            const int64_t o = (((int64_t)n * R + r) * 3) * HW + p;
            float x = lt[o], y = lt[o + HW], z = lt[o + 2 * HW];
            normalize3(x, y, z);
            m0 += x; m1 += y; m2 += z;
            if (chrom) { chrom[o] = x; chrom[o + HW] = y; chrom[o + 2 * HW] = z; }
        }
        m0 /= (float)R; m1 /= (float)R; m2 /= (float)R;
        normalize3(m0, m1, m2);
        if (chrom_mean) {
            chrom_mean[((int64_t)n * 3 + 0) * HW + p] = m0;
            chrom_mean[((int64_t)n * 3 + 1) * HW + p] = m1;
            chrom_mean[((int64_t)n * 3 + 2) * HW + p] = m2;
        }
        for (int r = 0; r < R; r++) {
            const int64_t o = (((int64_t)n * R + r) * 3) * HW + p;
            float x = lt[o], y = lt[o + HW], z = lt[o + 2 * HW];
            normalize3(x, y, z);
            const float d = (1.f - (x * m0 + y * m1 + z * m2)) * a * w;
            if (diff) diff[((int64_t)n * R + r) * HW + p] = d;
            dsum += d;
        }
        asum = a;
    }
    const float bs = block_sum(dsum, s_tmp);
    const float as = block_sum(asum, s_tmp);
    if (threadIdx.x == 0) { atomicAdd(&sums[0], (double)bs); atomicAdd(&sums[1], (double)as); }
}

// g_lt = gscale * dL/dlt, with L = sum(diff) / sum(alpha) / R  (gscale = upstream grad * loss weight)
__global__ void __launch_bounds__(128) chrom_bwd_kernel(const float* __restrict__ lt, const float* __restrict__ alpha,
                                                      const float* __restrict__ img, int R, int64_t HW, int N,
                                                      const double* __restrict__ sums, const float* __restrict__ gscale_ptr,
                                                      float gscale, float* __restrict__ g_lt, int accumulate) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW * N) return;
    const int n = (int)(pix / HW);
    const int64_t p = pix % HW;
    const float a = alpha[pix];
    float w = 1.f;
    if (img) {
        const float i0 = img[((int64_t)n * 3 + 0) * HW + p], i1 = img[((int64_t)n * 3 + 1) * HW + p], i2 = img[((int64_t)n * 3 + 2) * HW + p];
        w = fminf(sqrtf(i0 * i0 + i1 * i1 + i2 * i2) * 20.f, 1.0f);
    }
    const float gs = gscale * (gscale_ptr ? *gscale_ptr : 1.f);
    const float s = gs * a * w / ((float)sums[1] * (float)R);
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    for (int r = 0; r < R; r++) {
        const int64_t o = (((int64_t)n * R + r) * 3) * HW + p;
        float x = lt[o], y = lt[o + HW], z = lt[o + 2 * HW];
        normalize3(x, y, z);
        m0 += x; m1 += y; m2 += z;
    }
    m0 /= (float)R; m1 /= (float)R; m2 /= (float)R;
    normalize3(m0, m1, m2);
    for (int r = 0; r < R; r++) {
        const int64_t o = (((int64_t)n * R + r) * 3) * HW + p;
        const float lx = lt[o], ly = lt[o + HW], lz = lt[o + 2 * HW];
        const float nr = fmaxf(sqrtf(lx * lx + ly * ly + lz * lz), 1e-12f);
        const float x = lx / nr, y = ly / nr, z = lz / nr;
        const float cm = x * m0 + y * m1 + z * m2;
        const float gx = -s * (m0 - x * cm) / nr, gy = -s * (m1 - y * cm) / nr, gz = -s * (m2 - z * cm) / nr;
        if (accumulate) { g_lt[o] += gx; g_lt[o + HW] += gy; g_lt[o + 2 * HW] += gz; }
        else { g_lt[o] = gx; g_lt[o + HW] = gy; g_lt[o + 2 * HW] = gz; }
    }
}

// ---- cropped, alpha-masked L1 between out and gt:  loss = mean |out*a - gt*a| over the central crop -------
__global__ void __launch_bounds__(256) l1_masked_kernel(const float* __restrict__ out, const float* __restrict__ gt,
                                                      const float* __restrict__ alpha, int N, int Cc, int H, int W, int crop,
                                                      float weight, float* __restrict__ g_out, double* __restrict__ loss_sum) {
    __shared__ float s_tmp[8];
    const int64_t total = (int64_t)N * Cc * H * W;
    const double cnt = (double)N * Cc * (H - 2 * crop) * (W - 2 * crop);
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const int n = (int)(i / ((int64_t)W * H * Cc));
        float g = 0.f;
        if (h >= crop && h < H - crop && w >= crop && w < W - crop) {
            const float a = alpha[((int64_t)n * H + h) * W + w];
            const float d = out[i] * a - gt[i] * a;
            acc += fabsf(d);
            g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * a * (float)((double)weight / cnt);
        }
        if (g_out) g_out[i] = g;
    }
    const float bs = block_sum(acc, s_tmp);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, (double)bs * (double)weight / cnt);
}

// ---- Adam -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                 float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                 float bc1, float bc2_sqrt, float gscale) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = ((float4*)p)[i], gg = ((const float4*)g)[i], mm = ((float4*)m)[i], vv = ((float4*)v)[i];
        float* P = (float*)&pp; float* G = (float*)&gg; float* M = (float*)&mm; float* V = (float*)&vv;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float ge = G[e] * gscale;
            M[e] = b1 * M[e] + (1.f - b1) * ge;
            V[e] = b2 * V[e] + (1.f - b2) * ge * ge;
            const float denom = sqrtf(V[e]) / bc2_sqrt + eps;
            P[e] -= (lr / bc1) * (M[e] / denom);
        }
        ((float4*)p)[i] = pp; ((float4*)m)[i] = mm; ((float4*)v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = n4 << 2; i < n; i++) {
            const float ge = g[i] * gscale;
            m[i] = b1 * m[i] + (1.f - b1) * ge;
            v[i] = b2 * v[i] + (1.f - b2) * ge * ge;
            p[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
        }
    }
}

}  // namespace

extern "C" int rnr_chrom_loss_fwd(const float* rays_lt, const float* alpha, const float* img, int R, int N, int H, int W,
                                  float* chrom, float* chrom_mean, float* diff, double* sums, void* stream) {
    const int64_t HW = (int64_t)H * W;
    chrom_fwd_kernel<<<rnr_cdiv(HW * N, 128), 128, 0, (cudaStream_t)stream>>>(rays_lt, alpha, img, R, HW, N, chrom, chrom_mean, diff, sums);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_chrom_loss_bwd(const float* rays_lt, const float* alpha, const float* img, int R, int N, int H, int W,
                                  const double* sums, const float* gscale_dev, float gscale, float* g_lt, int accumulate,
                                  void* stream) {
    const int64_t HW = (int64_t)H * W;
    chrom_bwd_kernel<<<rnr_cdiv(HW * N, 128), 128, 0, (cudaStream_t)stream>>>(rays_lt, alpha, img, R, HW, N, sums, gscale_dev, gscale, g_lt, accumulate);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_l1_masked(const float* out, const float* gt, const float* alpha, int N, int C, int H, int W, int crop,
                             float weight, float* g_out, double* loss_sum, void* stream) {
    RNR_REQUIRE(H > 2 * crop && W > 2 * crop, "l1_masked: crop too large");
    const int64_t total = (int64_t)N * C * H * W;
    int blocks = rnr_cdiv(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    l1_masked_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, gt, alpha, N, C, H, W, crop, weight, g_out, loss_sum);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, int step, float gscale, void* stream) {
    RNR_REQUIRE(step >= 1, "adam: step must be >= 1");
    RNR_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam: buffers must be 16-byte aligned");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    int blocks = rnr_cdiv(n / 4 + 1, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), gscale);
    RNR_LAUNCH_CHECK();
    return 0;
}
