// The two "small" losses of the RNR iteration as CUDA kernels (they used to be ~100 ATen launches under autograd):
//   lighting L1      train_rnr.py:571-579   sum |l_init - reconstruct_sh(coeff, basis_val)| over covered / uncovered light samples,
//                                           each divided by its sample count and weighted
//   albedo mean      train_rnr.py:596-607   | mean over touched texels of flatten_mipmap(ch 3:6 / 0:3) - 0.5 |, "touched" = differs
//                                           from the initial flattened texture in any of the 3 channels
// Forward values are accumulated into device doubles; the gradients are produced in the layouts the existing scatter kernels
// take (rnr_sh_project for the coefficients, rnr_flatten_mipmap(backward) for the texture levels).
#include "common.cuh"

namespace {

__device__ __forceinline__ double block_sum_d(double v, double* s_tmp) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_tmp[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += s_tmp[i];
    return r;     // valid in thread 0
}

// one warp per light sample: est[c] = sum_b basis[s,b] coeff[b,c] (warp-shuffle inner product), d = init - est
//   loss += |d| * w_s ;  sgn[s,c] = d loss / d est = -sign(d) * w_s        (w_s = w_cov / n_cov or w_unc / n_unc)
__global__ void __launch_bounds__(256) lighting_l1_kernel(const float* __restrict__ basis, const float* __restrict__ coeff,
                                                        const float* __restrict__ l_init, const unsigned char* __restrict__ mask,
                                                        int S, int B, float w_cov, float w_unc, float* __restrict__ sgn,
                                                        double* __restrict__ loss) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double s_tmp[8];
    const int lane = threadIdx.x & 31;
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    double acc = 0.0;
    if (s < S) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const float* brow = basis + (int64_t)s * B;
        for (int b = lane; b < B; b += 32) {
            const float bv = brow[b];
            a0 += bv * coeff[b * 3 + 0]; a1 += bv * coeff[b * 3 + 1]; a2 += bv * coeff[b * 3 + 2];
        }
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            const float w = mask[s] ? w_cov : w_unc;
            const float est[3] = {a0, a1, a2};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float d = l_init[s * 3 + c] - est[c];
                acc += (double)fabsf(d) * (double)w;
                sgn[s * 3 + c] = (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f)) * w;
            }
        }
    }
    const double bs = block_sum_d(acc, s_tmp);
    if (threadIdx.x == 0 && bs != 0.0) atomicAdd(loss, bs);
}

// sums[0] = #touched (diffuse, ch 0:3), sums[1] = #touched (specular, ch 3:6), sums[2 + c] = sum of channel c over its group's touched texels
__global__ void __launch_bounds__(256) albedo_reduce_kernel(const float* __restrict__ tex6, const float* __restrict__ init6, int64_t P,
                                                          double* __restrict__ sums) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double s_tmp[8];
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        float t[6], o[6];
#pragma unroll
        for (int c = 0; c < 6; c++) { t[c] = tex6[i * 6 + c]; o[c] = init6[i * 6 + c]; }
        const bool vd = (t[0] != o[0]) || (t[1] != o[1]) || (t[2] != o[2]);
        const bool vs = (t[3] != o[3]) || (t[4] != o[4]) || (t[5] != o[5]);
        if (vd) { a[0] += 1.0; a[2] += t[0]; a[3] += t[1]; a[4] += t[2]; }
        if (vs) { a[1] += 1.0; a[5] += t[3]; a[6] += t[4]; a[7] += t[5]; }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double bs = block_sum_d(a[k], s_tmp);
        if (threadIdx.x == 0 && bs != 0.0) atomicAdd(sums + k, bs);
    }
}

// gout[p, c] = d (w_alb * loss_alb) / d tex6[p, c] = touched_group(p) * sign(mean_c - 0.5) / (3 * cnt_group) * w_alb; block 0 adds the
// loss value itself to *loss.  (The touched mask is a constant of the graph, exactly as in the reference's autograd.)
__global__ void __launch_bounds__(256) albedo_grad_kernel(const float* __restrict__ tex6, const float* __restrict__ init6, int64_t P,
                                                        const double* __restrict__ sums, float w_alb, float* __restrict__ gout,
                                                        double* __restrict__ loss) {
    pdl_launch_dependents();
    pdl_wait();
    float gc[6];
    double l = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const double cnt = sums[c < 3 ? 0 : 1];
        if (cnt > 0.0) {
            const double mean = sums[2 + c] / cnt - 0.5;
            l += fabs(mean) / 3.0;
            gc[c] = (float)((mean > 0.0 ? 1.0 : (mean < 0.0 ? -1.0 : 0.0)) / (3.0 * cnt) * (double)w_alb);
        } else gc[c] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) atomicAdd(loss, l * (double)w_alb);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        float t[6], o[6];
#pragma unroll
        for (int c = 0; c < 6; c++) { t[c] = tex6[i * 6 + c]; o[c] = init6[i * 6 + c]; }
        const bool vd = (t[0] != o[0]) || (t[1] != o[1]) || (t[2] != o[2]);
        const bool vs = (t[3] != o[3]) || (t[4] != o[4]) || (t[5] != o[5]);
#pragma unroll
        for (int c = 0; c < 6; c++) gout[i * 6 + c] = ((c < 3) ? vd : vs) ? gc[c] : 0.f;
    }
}

}  // namespace

extern "C" int rnr_lighting_l1(const float* basis, const float* coeff, const float* l_init, const unsigned char* mask, int S, int B,
                               float w_cov, float w_unc, float* sgn, double* loss, void* stream) {
    RNR_REQUIRE(basis && coeff && l_init && mask && sgn && loss && S >= 1 && B >= 1, "rnr_lighting_l1: bad arguments");
    RNR_PDL_LAUNCH(lighting_l1_kernel, rnr_cdiv((int64_t)S * 32, 256), 256, 0, stream, basis, coeff, l_init, mask, S, B, w_cov, w_unc, sgn, loss);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_albedo_mean_loss(const float* tex6, const float* init6, int64_t P, float w_alb, double* sums, float* gout, double* loss,
                                    void* stream) {
    RNR_REQUIRE(tex6 && init6 && sums && gout && loss && P >= 1, "rnr_albedo_mean_loss: bad arguments");
    int blocks = rnr_cdiv(P, 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    RNR_PDL_LAUNCH(albedo_reduce_kernel, blocks, 256, 0, stream, tex6, init6, P, sums);
    RNR_LAUNCH_CHECK();
    RNR_PDL_LAUNCH(albedo_grad_kernel, blocks, 256, 0, stream, tex6, init6, P, sums, w_alb, gout, loss);
    RNR_LAUNCH_CHECK();
    return 0;
}
