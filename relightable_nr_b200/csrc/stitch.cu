// Light-probe stitching: back-projection of the image background of one view into an equirectangular
// environment map.  Reference: stitch_lp.py:27-35 (camera2ray), :22-24 (spherical_mapping), :136-147 (per-view scatter).
//
// The reference scatters with numpy fancy indexing,  env[v, u] += img[mask];  count[v, u] += 1,  which is NOT an accumulation:
// when several background pixels of a view fall into the same probe texel only the LAST one in row-major pixel order is added,
// and the texel's count goes up by one per view.  The device version keeps exactly that rule:
//   pass 1  every background pixel computes its texel in fp64 (same operation order as the numpy code, round-half-even like
//           np.round) and bids for it with atomicMax(pixel index);
//   pass 2  the winning pixel of each texel adds its colour (fp64 accumulator, like the reference's float64 `env`), bumps the
//           count and re-arms the bid for the next view.
// Cold path: one launch pair per view, HBM-trivial (a 512^2 view touches 3 MB).
#include "common.cuh"
#include <math.h>

namespace {

struct StitchCam {
    double kinv[9];      // inverse intrinsics, row-major
    double rinv[9];      // inverse of the rotation part of the world->camera pose
};

__global__ void __launch_bounds__(256) stitch_bid_kernel(const unsigned char* __restrict__ bg, const StitchCam cam, int h, int w,
                                                         int lp_h, int lp_w, int* __restrict__ texel, int* __restrict__ winner) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    int t = -1;
    if (bg[p]) {
        const int y = p / w, x = p - y * w;
        // camera2ray: K^-1 (x + .5, y + .5, 1), then R^-1, then normalise   (plain multiply-adds in the reference's order; no FMA
        // contraction so that the fp64 result matches numpy's dot bit for bit wherever the libm calls below agree)
        const double px = (double)x + 0.5, py = (double)y + 0.5;
        double c[3], d[3];
#pragma unroll
        for (int i = 0; i < 3; i++) c[i] = __dadd_rn(__dadd_rn(__dmul_rn(cam.kinv[i * 3], px), __dmul_rn(cam.kinv[i * 3 + 1], py)), cam.kinv[i * 3 + 2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
            d[i] = __dadd_rn(__dadd_rn(__dmul_rn(cam.rinv[i * 3], c[0]), __dmul_rn(cam.rinv[i * 3 + 1], c[1])), __dmul_rn(cam.rinv[i * 3 + 2], c[2]));
        const double len = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
        d[0] /= len; d[1] /= len; d[2] /= len;
        // spherical_mapping: u = atan2(z, x) / 2pi + 1/2,  v = acos(y) / pi;  scaled, clipped from above, rounded half-to-even
        const double kPi = 3.141592653589793;
        double u = __dadd_rn(__dmul_rn(__dmul_rn(atan2(d[2], d[0]), 0.5) / kPi, 1.0), 0.5);
        double v = __dmul_rn(acos(d[1]), 1.0) / kPi;
        u = fmin(__dmul_rn(u, (double)lp_w), (double)lp_w - 1.0);
        v = fmin(__dmul_rn(v, (double)lp_h), (double)lp_h - 1.0);
        const int iu = (int)rint(u), iv = (int)rint(v);
        if (iu >= 0 && iu < lp_w && iv >= 0 && iv < lp_h) {       // (NaN directions -- a degenerate camera -- fall out here)
            t = iv * lp_w + iu;
            atomicMax(winner + t, p);
        }
    }
    texel[p] = t;
}

__global__ void __launch_bounds__(256) stitch_add_kernel(const float* __restrict__ img, const int* __restrict__ texel, int npix,
                                                         int* __restrict__ winner, double* __restrict__ env, float* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int t = texel[p];
    if (t < 0 || winner[t] != p) return;
    // the one winner of texel t: nobody else touches env / count / winner at t in this launch
#pragma unroll
    for (int c = 0; c < 3; c++) {
        env[(int64_t)t * 3 + c] += (double)img[(int64_t)p * 3 + c];
        count[(int64_t)t * 3 + c] += 1.f;
    }
    winner[t] = -1;
}

__global__ void __launch_bounds__(256) stitch_finish_kernel(double* __restrict__ env, const float* __restrict__ count,
                                                            unsigned char* __restrict__ mask, int64_t ntex) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntex) return;
    const float c0 = count[t * 3], c1 = count[t * 3 + 1], c2 = count[t * 3 + 2];
    const bool hit = (c0 + c1 + c2) > 0.f;                       // stitch_lp.py:149-150
    if (hit) { env[t * 3] /= (double)c0; env[t * 3 + 1] /= (double)c1; env[t * 3 + 2] /= (double)c2; }
    mask[t] = hit ? 255 : 0;
}

}  // namespace

// One view of stitch_lp.py:136-147.  img [h, w, 3] fp32 (channel order as read), bg [h, w] uint8 (non-zero = background pixel to
// project), kinv / rinv: 9 doubles each on the HOST.  texel [h*w] int32 scratch; winner [lp_h*lp_w] int32, all -1 before the first
// view (left all -1 again); env [lp_h, lp_w, 3] fp64 and count [lp_h, lp_w, 3] fp32 accumulate across views.
extern "C" int rnr_stitch_view(const float* img, const unsigned char* bg, const double* kinv, const double* rinv, int h, int w,
                               int lp_h, int lp_w, int* texel, int* winner, double* env, float* count, void* stream) {
    RNR_REQUIRE(img && bg && kinv && rinv && texel && winner && env && count, "rnr_stitch_view: null pointer");
    RNR_REQUIRE(h > 0 && w > 0 && lp_h > 0 && lp_w > 0 && (int64_t)h * w < (1ll << 31) && (int64_t)lp_h * lp_w < (1ll << 31),
                "rnr_stitch_view: bad sizes %dx%d -> %dx%d", h, w, lp_h, lp_w);
    StitchCam cam;
    for (int i = 0; i < 9; i++) { cam.kinv[i] = kinv[i]; cam.rinv[i] = rinv[i]; }
    const int npix = h * w;
    stitch_bid_kernel<<<rnr_cdiv(npix, 256), 256, 0, (cudaStream_t)stream>>>(bg, cam, h, w, lp_h, lp_w, texel, winner);
    RNR_LAUNCH_CHECK();
    stitch_add_kernel<<<rnr_cdiv(npix, 256), 256, 0, (cudaStream_t)stream>>>(img, texel, npix, winner, env, count);
    RNR_LAUNCH_CHECK();
    return 0;
}

// stitch_lp.py:149-150: env /= count where any view hit the texel; mask [lp_h, lp_w] uint8 = 255 there.
extern "C" int rnr_stitch_finish(double* env, const float* count, unsigned char* mask, int lp_h, int lp_w, void* stream) {
    RNR_REQUIRE(env && count && mask && lp_h > 0 && lp_w > 0, "rnr_stitch_finish: bad arguments");
    const int64_t ntex = (int64_t)lp_h * lp_w;
    stitch_finish_kernel<<<(unsigned)rnr_cdiv(ntex, 256), 256, 0, (cudaStream_t)stream>>>(env, count, mask, ntex);
    RNR_LAUNCH_CHECK();
    return 0;
}
