// Per-pixel maps derived from the G-buffer.
//   camera.get_view_dir_map      camera.py:5-32     (-K^-1 [u+.5, v+.5, 1] normalised, rotated by R^-1)
//   render.get_TBN_map           render.py:124-168  (per-face tangent from edge/UV deltas, Gram-Schmidt per pixel)
//   render.interp_vertex_attr    render.py:11-28    (barycentric blend of per-vertex attributes)
// All HBM-bound, one thread per pixel, channel-interleaved outputs staged so that stores are coalesced.
#include "pixel.cuh"

namespace {

__global__ void __launch_bounds__(256) view_dir_kernel(const float* __restrict__ proj_inv, const float* __restrict__ R_inv,
                                                     float* __restrict__ out_world, float* __restrict__ out_cam, int N, int H, int W) {
    __shared__ float s_w[256 * 3], s_c[256 * 3];
    const int64_t P = (int64_t)H * W;
    const int n = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * 256, pix = p0 + threadIdx.x;
    if (pix < P) {
        const float* Ki = proj_inv + n * 9;
        const float* Ri = R_inv + n * 9;
        const float u = (float)(pix % W) + 0.5f, v = (float)(pix / W) + 0.5f;
        float x = -(Ki[0] * u + Ki[1] * v + Ki[2]);
        float y = -(Ki[3] * u + Ki[4] * v + Ki[5]);
        float z = -(Ki[6] * u + Ki[7] * v + Ki[8]);
        normalize3(x, y, z);
        s_c[threadIdx.x * 3 + 0] = x; s_c[threadIdx.x * 3 + 1] = y; s_c[threadIdx.x * 3 + 2] = z;
        float wx = Ri[0] * x + Ri[1] * y + Ri[2] * z;
        float wy = Ri[3] * x + Ri[4] * y + Ri[5] * z;
        float wz = Ri[6] * x + Ri[7] * y + Ri[8] * z;
        normalize3(wx, wy, wz);
        s_w[threadIdx.x * 3 + 0] = wx; s_w[threadIdx.x * 3 + 1] = wy; s_w[threadIdx.x * 3 + 2] = wz;
    }
    __syncthreads();
    const int64_t np = (P - p0) < 256 ? (P - p0) : 256;
    for (int64_t i = threadIdx.x; i < np * 3; i += 256) {
        out_world[((int64_t)n * P + p0) * 3 + i] = s_w[i];
        if (out_cam) out_cam[((int64_t)n * P + p0) * 3 + i] = s_c[i];
    }
}

// per-face tangent (render.py:139-148); flag[0] |= 1 when a NaN is produced
__global__ void __launch_bounds__(256) face_tangent_kernel(const float* __restrict__ fv, const float* __restrict__ fvt,
                                                         float* __restrict__ tangent, int* __restrict__ flag, int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const float* v = fv + (int64_t)f * 9;
    const float* t = fvt + (int64_t)f * 6;
    const float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
    const float e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
    const float du1 = t[2] - t[0], dv1 = t[3] - t[1], du2 = t[4] - t[0], dv2 = t[5] - t[1];
    const float det = du1 * dv2 - du2 * dv1;
    const float fi = 1.0f / fmaxf(det, 1e-8f);
    float tx = fi * (dv2 * e1x - dv1 * e2x), ty = fi * (dv2 * e1y - dv1 * e2y), tz = fi * (dv2 * e1z - dv1 * e2z);
    if (tx != tx || ty != ty || tz != tz || det != det) atomicOr(flag, 1);
    normalize3(tx, ty, tz);
    tangent[(int64_t)f * 3 + 0] = tx; tangent[(int64_t)f * 3 + 1] = ty; tangent[(int64_t)f * 3 + 2] = tz;
}

// TBN [P,3,3] with columns (T, B, N); face index -1 (background) addresses the last face like the reference's
// negative indexing -- with a zero normal the whole matrix comes out 0 there.
__global__ void __launch_bounds__(128) tbn_kernel(const float* __restrict__ normal, const int* __restrict__ fidx,
                                                const float* __restrict__ tangent, float* __restrict__ tbn, int* __restrict__ flag,
                                                int64_t P, int nf) {
    __shared__ float s[128 * 9];
    const int64_t p0 = (int64_t)blockIdx.x * 128, pix = p0 + threadIdx.x;
    if (pix < P) {
        int f = fidx[pix];
        if (f < 0) f += nf;
        f = min(max(f, 0), nf - 1);
        float tx = tangent[(int64_t)f * 3], ty = tangent[(int64_t)f * 3 + 1], tz = tangent[(int64_t)f * 3 + 2];
        float nx = normal[pix * 3], ny = normal[pix * 3 + 1], nz = normal[pix * 3 + 2];
        normalize3(nx, ny, nz);
        float bx = ny * tz - nz * ty, by = nz * tx - nx * tz, bz = nx * ty - ny * tx;
        normalize3(bx, by, bz);
        tx = by * nz - bz * ny; ty = bz * nx - bx * nz; tz = bx * ny - by * nx;
        normalize3(tx, ty, tz);
        float* o = s + threadIdx.x * 9;
        o[0] = tx; o[1] = bx; o[2] = nx;
        o[3] = ty; o[4] = by; o[5] = ny;
        o[6] = tz; o[7] = bz; o[8] = nz;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 9; i++) bad |= (o[i] != o[i]);
        if (bad) atomicOr(flag, 1);
    }
    __syncthreads();
    const int64_t np = (P - p0) < 128 ? (P - p0) : 128;
    for (int64_t i = threadIdx.x; i < np * 9; i += 128) tbn[p0 * 9 + i] = s[i];
}

// out[n,pix,a] = sum_k w[n,pix,k] * attr[nb, faces[nf_b, f, k], a]; attr batch 1 or N; A <= 16
__global__ void __launch_bounds__(256) interp_attr_kernel(const float* __restrict__ attr, int attr_batch, int nv, int A,
                                                        const int* __restrict__ faces, int nf, const int* __restrict__ fidx,
                                                        const float* __restrict__ w, float* __restrict__ out, int64_t P) {
    const int n = blockIdx.y;
    const int64_t pix = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (pix >= P) return;
    int f = fidx[(int64_t)n * P + pix];
    if (f < 0) f += nf;
    f = min(max(f, 0), nf - 1);
    const int* fv = faces + ((int64_t)n * nf + f) * 3;
    const float* wp = w + ((int64_t)n * P + pix) * 3;
    const float* ab = attr + (attr_batch == 1 ? 0 : (int64_t)n * nv * A);
    const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const float* a0 = ab + (int64_t)fv[0] * A;
    const float* a1 = ab + (int64_t)fv[1] * A;
    const float* a2 = ab + (int64_t)fv[2] * A;
    float* o = out + ((int64_t)n * P + pix) * A;
    for (int a = 0; a < A; a++) o[a] = a0[a] * w0 + a1[a] * w1 + a2[a] * w2;
}

}  // namespace

extern "C" int rnr_view_dir_map(const float* proj_inv, const float* R_inv, float* out_world, float* out_cam, int N, int H, int W,
                                void* stream) {
    if ((int64_t)N * H * W == 0) return 0;
    dim3 grid(rnr_cdiv((int64_t)H * W, 256), N);
    view_dir_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(proj_inv, R_inv, out_world, out_cam, N, H, W);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_face_tangents(const float* faces_v, const float* faces_vt, float* tangent, int* nan_flag, int nf, void* stream) {
    if (nf == 0) return 0;
    face_tangent_kernel<<<rnr_cdiv(nf, 256), 256, 0, (cudaStream_t)stream>>>(faces_v, faces_vt, tangent, nan_flag, nf);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_tbn_map(const float* normal_map, const int* face_index_map, const float* tangent, float* tbn, int* nan_flag,
                           int64_t P, int nf, void* stream) {
    if (P == 0) return 0;
    RNR_REQUIRE(nf > 0, "rnr_tbn_map: empty mesh");
    tbn_kernel<<<rnr_cdiv(P, 128), 128, 0, (cudaStream_t)stream>>>(normal_map, face_index_map, tangent, tbn, nan_flag, P, nf);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_interp_vertex_attr(const float* attr, int attr_batch, int nv, int A, const int* faces, int nf,
                                      const int* face_index_map, const float* weight_map, float* out, int N, int64_t P, void* stream) {
    if ((int64_t)N * P == 0) return 0;
    RNR_REQUIRE(nf > 0 && nv > 0, "rnr_interp_vertex_attr: empty mesh");
    dim3 grid(rnr_cdiv(P, 256), N);
    interp_attr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(attr, attr_batch, nv, A, faces, nf, face_index_map, weight_map, out, P);
    RNR_LAUNCH_CHECK();
    return 0;
}
