// Validation metrics on the device: the masked MAE / MSE / PSNR family of metric.py:19-84, which the
// reference computes per view in numpy after a GPU -> CPU copy of the rendered image, the ground truth and the mask
// (train_rnr.py:626-633 every iteration, :707-887 at validation).  Two kernels per batch, one read-back of 8 doubles per image:
//   metric_bbox_kernel   bounding box of mask == 1 and its pixel count (metric.py:47-52)
//   metric_sums_kernel   sum |d| and sum d^2 over the image and over the bounding box, d = est*[m==1] - gt*[m==1] (metric.py:33-34,60-71)
#include "common.cuh"

namespace {

// box[n] = {xmin, xmax, ymin, ymax} (initialised to {W, -1, H, -1}), cnt[n] = number of mask == 1 pixels
__global__ void __launch_bounds__(256) metric_bbox_kernel(const float* __restrict__ mask, int N, int H, int W, int* __restrict__ box,
                                                        unsigned long long* __restrict__ cnt) {
    const int n = blockIdx.y;
    int xmin = W, xmax = -1, ymin = H, ymax = -1;
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)H * W; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask[(int64_t)n * H * W + i] == 1.f) {
            const int x = (int)(i % W), y = (int)(i / W);
            xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
            c++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0 && c > 0) {
        atomicMin(box + n * 4 + 0, xmin); atomicMax(box + n * 4 + 1, xmax);
        atomicMin(box + n * 4 + 2, ymin); atomicMax(box + n * 4 + 3, ymax);
        atomicAdd(cnt + n, c);
    }
}

// sums[n] = {sum |d|, sum d^2, sum |d| inside the box, sum d^2 inside the box}; est / gt [N,C,H,W]
__global__ void __launch_bounds__(256) metric_sums_kernel(const float* __restrict__ est, const float* __restrict__ gt, const float* __restrict__ mask,
                                                        int N, int C, int H, int W, const int* __restrict__ box, double* __restrict__ sums) {
    const int n = blockIdx.y;
    const int xmin = box[n * 4], xmax = box[n * 4 + 1], ymin = box[n * 4 + 2], ymax = box[n * 4 + 3];
    double a[4] = {0, 0, 0, 0};
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i % HW;
        if (mask[(int64_t)n * HW + p] != 1.f) continue;          // both images are zeroed where the mask is not 1: d = 0
        const float d = fabsf(est[(int64_t)n * C * HW + i] - gt[(int64_t)n * C * HW + i]);
        a[0] += d; a[1] += (double)d * d;
        const int x = (int)(p % W), y = (int)(p / W);
        if (x >= xmin && x <= xmax && y >= ymin && y <= ymax) { a[2] += d; a[3] += (double)d * d; }
    }
    __shared__ double s_tmp[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double v = a[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += s_tmp[w];
            if (r != 0.0) atomicAdd(sums + n * 4 + k, r);
        }
    }
}

}  // namespace

extern "C" int rnr_metric_sums(const float* est, const float* gt, const float* mask, int N, int C, int H, int W, int* box,
                               unsigned long long* cnt, double* sums, void* stream) {
    RNR_REQUIRE(est && gt && mask && box && cnt && sums && N >= 1 && C >= 1, "rnr_metric_sums: bad arguments");
    int blocks = rnr_cdiv((int64_t)H * W, 256);
    if (blocks > 148 * 2) blocks = 148 * 2;
    dim3 grid(blocks, N);
    metric_bbox_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, N, H, W, box, cnt);
    RNR_LAUNCH_CHECK();
    metric_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(est, gt, mask, N, C, H, W, box, sums);
    RNR_LAUNCH_CHECK();
    return 0;
}
