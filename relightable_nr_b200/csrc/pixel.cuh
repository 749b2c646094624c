// Per-pixel device primitives shared by the module-level kernels and the fused producers.
#pragma once
#include "common.cuh"

// misc.interpolate_bilinear (misc.py:5-42): 4-tap gather with hard validity mask and the
// right/bottom edge weight fix-up.  Returns tap indices (into an [Hd, Wd] grid) and weights.
struct Bilin {
    int i00, i10, i01, i11;     // (y0,x0) (y1,x0) (y0,x1) (y1,x1) as y*Wd + x
    float w00, w10, w01, w11;
};

__device__ __forceinline__ Bilin bilinear_setup(float x, float y, int Wd, int Hd) {
    Bilin b;
    const float m = (x >= 0.f && x <= (float)(Wd - 1) && y >= 0.f && y <= (float)(Hd - 1)) ? 1.f : 0.f;
    // NaN / huge coordinates: mask is 0; keep indices in range
    float fx = floorf(x), fy = floorf(y);
    if (!(fx >= -1.f)) fx = -1.f;
    if (!(fy >= -1.f)) fy = -1.f;
    if (fx > (float)Wd) fx = (float)Wd;
    if (fy > (float)Hd) fy = (float)Hd;
    int x0 = (int)fx, y0 = (int)fy;
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), Wd - 1); x1 = min(max(x1, 0), Wd - 1);
    y0 = min(max(y0, 0), Hd - 1); y1 = min(max(y1, 0), Hd - 1);
    b.i00 = y0 * Wd + x0; b.i10 = y1 * Wd + x0; b.i01 = y0 * Wd + x1; b.i11 = y1 * Wd + x1;
    const int x0w = x0 - (x0 == x1 ? 1 : 0);
    const int y0w = y0 - (y0 == y1 ? 1 : 0);
    const float ax = (float)x1 - x, bx = x - (float)x0w;
    const float ay = (float)y1 - y, by = y - (float)y0w;
    b.w00 = ax * ay * m; b.w10 = ax * by * m; b.w01 = bx * ay * m; b.w11 = bx * by * m;
    return b;
}

// F.normalize(v, dim) for a 3-vector: v / max(||v||, 1e-12)
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    x /= n; y /= n; z /= n;
}

#define RNR_PI_F 3.14159265358979323846f
// render.spherical_mapping (render.py:96-102): equirect uv of a direction
__device__ __forceinline__ void spherical_uv(float x, float y, float z, float& u, float& v) {
    u = atan2f(z, x) * 0.5f / RNR_PI_F + 0.5f;
    v = acosf(y) * 1.0f / RNR_PI_F;
}
