// Dense EdgeConv of gcn_lib (network.DenseDeepGCN, network.py:256-315):
//   gcn_lib/dense/torch_vertex.py:23-35  EdgeConv4D:  max_k  nn(cat[x_i, x_j - x_i])
//   gcn_lib/dense/torch_nn.py:55-64      BasicConv:   Conv2d(1x1) -> activation -> BatchNorm2d   (activation BEFORE the norm)
// The 1x1 convolution is linear in (x_i, x_j):  W [x_i ; x_j - x_i] + b = (W1 - W2) x_i + W2 x_j + b = P[i] + Q[j], so the
// [V, k, 2C] edge tensor of the reference is never built: one GEMM gives P | Q per VERTEX, and this kernel walks the k
// neighbours of a vertex (one warp per vertex, lanes over channels, coalesced 128-byte gathers of Q rows), applies the
// activation, and keeps max / min / sum / sum^2.  BatchNorm with batch statistics is a per-channel affine map a*s + t applied
// after the activation, so   max_k (a_k s + t) = s max_k a_k + t  (s >= 0)  or  s min_k a_k + t  (s < 0): the finish kernel
// picks the right extremum once the statistics (fp64 sums over all V*k edges) are known.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) edgeconv_reduce_kernel(const float* __restrict__ pq, const int* __restrict__ nbr, int V, int K, int C,
                                                            float slope, float* __restrict__ amax, float* __restrict__ amin,
                                                            double* __restrict__ sums) {
    extern __shared__ float s_red[];          // [8][2C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = gridDim.x * 8;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float s1 = 0.f, s2 = 0.f;
        for (int v = blockIdx.x * 8 + warp; v < V; v += nwarps) {
            float mx = -INFINITY, mn = INFINITY;
            if (c < C) {
                const float p = pq[(int64_t)v * 2 * C + c];
                for (int k = 0; k < K; k++) {
                    const int j = nbr[(int64_t)v * K + k];
                    const float y = p + __ldg(pq + (int64_t)j * 2 * C + C + c);
                    const float a = y > 0.f ? y : y * slope;
                    mx = fmaxf(mx, a); mn = fminf(mn, a);
                    s1 += a; s2 += a * a;
                }
                amax[(int64_t)v * C + c] = mx;
                amin[(int64_t)v * C + c] = mn;
            }
        }
        if (c < C) { s_red[warp * 2 * C + c] = s1; s_red[warp * 2 * C + C + c] = s2; }
    }
    __syncthreads();
    if (sums)
        for (int i = threadIdx.x; i < 2 * C; i += 256) {
            double a = 0.0;
            for (int w = 0; w < 8; w++) a += (double)s_red[w * 2 * C + i];
            atomicAdd(sums + i, a);
        }
}

__global__ void __launch_bounds__(256) edgeconv_finish_kernel(const float* __restrict__ amax, const float* __restrict__ amin,
                                                            const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* running_mean,
                                                            float* running_var, float momentum, int training, int has_bn,
                                                            const float* __restrict__ residual, float* __restrict__ out, int V, int C) {
    const int64_t total = (int64_t)V * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float r;
        if (has_bn) {
            double m, var;
            if (training) {
                m = sums[c] / count;
                var = sums[C + c] / count - m * m;
                if (var < 0.0) var = 0.0;
            } else {
                m = (double)running_mean[c];
                var = (double)running_var[c];
            }
            const float sc = gamma[c] * (float)(1.0 / sqrt(var + (double)eps));
            const float sh = beta[c] - (float)m * sc;
            r = (sc >= 0.f ? amax[i] : amin[i]) * sc + sh;
        } else {
            r = amax[i];
        }
        if (residual) r += residual[i];
        out[i] = r;
    }
}

__global__ void edgeconv_running_kernel(const double* __restrict__ sums, double count, float* running_mean, float* running_var,
                                        float momentum, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
}

}  // namespace

extern "C" int rnr_edgeconv_reduce(const float* pq, const int32_t* nbr, int V, int K, int C, float slope, float* amax, float* amin,
                                   double* sums, void* stream) {
    RNR_REQUIRE(pq && nbr && amax && amin, "edgeconv: null argument");
    RNR_REQUIRE(V >= 1 && K >= 1 && C >= 1 && C <= 4096, "edgeconv: bad sizes V=%d K=%d C=%d", V, K, C);
    int blocks = rnr_cdiv(V, 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    edgeconv_reduce_kernel<<<blocks, 256, (size_t)8 * 2 * C * sizeof(float), (cudaStream_t)stream>>>(pq, nbr, V, K, C, slope, amax, amin, sums);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_edgeconv_finish(const float* amax, const float* amin, const double* sums, double count, const float* gamma,
                                   const float* beta, float eps, float* running_mean, float* running_var, float momentum,
                                   int training, const float* residual, float* out, int V, int C, void* stream) {
    const int has_bn = gamma != nullptr;
    RNR_REQUIRE(!has_bn || (beta && (training ? sums != nullptr : (running_mean && running_var))), "edgeconv finish: BatchNorm inputs missing");
    int blocks = rnr_cdiv((int64_t)V * C, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    edgeconv_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(amax, amin, sums, count, gamma, beta, eps, running_mean, running_var,
                                                                    momentum, training, has_bn, residual, out, V, C);
    RNR_LAUNCH_CHECK();
    if (has_bn && training && running_mean && running_var) {
        edgeconv_running_kernel<<<rnr_cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, running_mean, running_var, momentum, C);
        RNR_LAUNCH_CHECK();
    }
    return 0;
}
