// Proxy-mesh rasterizer: projection -> per-face set-up -> tile-culled z-buffer with fused G-buffer attributes.
//   nr.projection                        neural_renderer/projection.py:6-53
//   nr.vertices_to_faces + kernel 1      neural_renderer/vertices_to_faces.py:4-25, cuda/rasterize_cuda_kernel.cu:24-68
//   forward_face_index_map kernel 2      cuda/rasterize_cuda_kernel.cu:70-169
//   rasterize_rgbad vertical flip        neural_renderer/rasterize.py:313-321   (folded into the store address)
//   network.Rasterizer.forward           network.py:176-214  (perspective-correct weights, uv / normal / position maps)
//
// The reference tests every pixel against every face (O(P nf)).  Here one CTA owns a 32x8 pixel tile: it streams the
// packed per-face screen bounding boxes (8 B/face) once, keeps the faces whose box touches the tile by an ORDER-PRESERVING
// ballot compaction (ascending face index, so depth ties resolve to the lowest face index exactly as the reference's
// sequential loop does), stages their records through shared memory and lets each thread z-test its own pixel.
// No atomics, no per-tile lists in HBM, deterministic.  All arithmetic of the coverage / barycentric / depth test follows
// the reference's operation order in fp32 without FMA contraction (and its double-promoted sub-expressions in fp64), so the
// integer face-index map is bit-reproducible against the CPU oracle (oracle/raster.py).
// HBM traffic per view: 8 B/face x tiles (L2 resident) + 72 B per surviving face + the G-buffer stores, which are staged
// through shared memory so that every map is written as full-width contiguous rows of the tile.
#include "pixel.cuh"

namespace {

constexpr int TW = 32, TH = 8, NT = TW * TH;      // tile = 32 x 8 pixels, one thread per pixel
constexpr int CHUNK = 4;                          // faces culled per thread per pass (pass = 1024 faces)
constexpr int BATCH = 128;                        // surviving faces staged per shared-memory batch

// ---------------------------------------------------------------------------------------------
// projection.py:6-53
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ v, int v_batch, int nv, const float* __restrict__ K,
                                                    const float* __restrict__ R, const float* __restrict__ t,
                                                    const float* __restrict__ dist, const float* __restrict__ offset,
                                                    const float* __restrict__ scale, float orig_size, float eps,
                                                    float* __restrict__ out, int N) {
    const int n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const float* p = v + ((int64_t)(v_batch == 1 ? 0 : n) * nv + i) * 3;
    const float* Rn = R + n * 9;
    const float* Kn = K + n * 9;
    const float px = p[0], py = p[1], pz = p[2];
    const float x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, Rn[0]), __fmul_rn(py, Rn[1])), __fmul_rn(pz, Rn[2])), t[n * 3 + 0]);
    const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, Rn[3]), __fmul_rn(py, Rn[4])), __fmul_rn(pz, Rn[5])), t[n * 3 + 1]);
    const float z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, Rn[6]), __fmul_rn(py, Rn[7])), __fmul_rn(pz, Rn[8])), t[n * 3 + 2]);
    const float x_ = __fdiv_rn(x, __fadd_rn(z, eps)), y_ = __fdiv_rn(y, __fadd_rn(z, eps));
    float x__ = x_, y__ = y_;
    if (dist) {
        const float k1 = dist[n * 5 + 0], k2 = dist[n * 5 + 1], p1 = dist[n * 5 + 2], p2 = dist[n * 5 + 3], k3 = dist[n * 5 + 4];
        const float r = sqrtf(x_ * x_ + y_ * y_);
        const float r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
        const float rad = 1.f + k1 * r2 + k2 * r4 + k3 * r6;
        x__ = x_ * rad + 2.f * p1 * x_ * y_ + p2 * (r2 + 2.f * x_ * x_);
        y__ = y_ * rad + p1 * (r2 + 2.f * y_ * y_) + 2.f * p2 * x_ * y_;
    }
    float u = __fadd_rn(__fadd_rn(__fmul_rn(x__, Kn[0]), __fmul_rn(y__, Kn[1])), Kn[2]);
    float w = __fadd_rn(__fadd_rn(__fmul_rn(x__, Kn[3]), __fmul_rn(y__, Kn[4])), Kn[5]);
    if (offset && scale) {
        u = __fmul_rn(__fadd_rn(u, offset[n * 2 + 1]), scale[n * 2 + 1]);
        w = __fmul_rn(__fadd_rn(w, offset[n * 2 + 0]), scale[n * 2 + 0]);
    }
    w = __fadd_rn(orig_size, -w);
    const float half = __fdiv_rn(orig_size, 2.f);
    u = __fdiv_rn(__fmul_rn(2.f, __fadd_rn(u, -half)), orig_size);
    w = __fdiv_rn(__fmul_rn(2.f, __fadd_rn(w, -half)), orig_size);
    float* o = out + ((int64_t)n * nv + i) * 3;
    o[0] = u; o[1] = w; o[2] = z;
}

// ---------------------------------------------------------------------------------------------
// per-face set-up: gather (optional), back-face flag, inverse barycentric matrix, packed pixel bbox
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool backside(const float* f) {
    // rasterize_cuda_kernel.cu:40, :109
    return __fmul_rn(__fadd_rn(f[7], -f[1]), __fadd_rn(f[3], -f[0])) < __fmul_rn(__fadd_rn(f[4], -f[1]), __fadd_rn(f[6], -f[0]));
}

__global__ void __launch_bounds__(256) face_setup_kernel(const float* __restrict__ uvz, int nv, const int32_t* __restrict__ fidx,
                                                       int f_batch, const float* __restrict__ faces_in, int nf, int is,
                                                       float* __restrict__ faces_out, float* __restrict__ faces_inv,
                                                       int2* __restrict__ bbox, int N) {
    const int n = blockIdx.y;
    const int fi = blockIdx.x * blockDim.x + threadIdx.x;
    if (fi >= nf) return;
    const int64_t fo = ((int64_t)n * nf + fi);
    float f[9];
    if (faces_in) {
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = faces_in[fo * 9 + k];
    } else {
        const int32_t* id = fidx + ((int64_t)(f_batch == 1 ? 0 : n) * nf + fi) * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int vi = min(max(id[k], 0), nv - 1);
            const float* s = uvz + ((int64_t)n * nv + vi) * 3;
            f[3 * k] = s[0]; f[3 * k + 1] = s[1]; f[3 * k + 2] = s[2];
        }
        if (faces_out) {
#pragma unroll
            for (int k = 0; k < 9; k++) faces_out[fo * 9 + k] = f[k];
        }
    }
    float inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int2 bb = make_int2(1, 0);         // empty: x0 > x1
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 9; k++) finite &= (fabsf(f[k]) <= 3.0e38f);
    if (finite && !backside(f)) {
        // p[num][dim] = 0.5 * (face * is + is - 1)   (:44-49; the double 0.5 is an exact scaling)
        float p[3][2];
        const float fis = (float)is;
#pragma unroll
        for (int num = 0; num < 3; num++)
#pragma unroll
            for (int d = 0; d < 2; d++)
                p[num][d] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fmul_rn(f[3 * num + d], fis), fis), -1.f));
        auto mm = [](float a, float b, float c, float d) { return __fadd_rn(__fmul_rn(a, b), -__fmul_rn(c, d)); };
        inv[0] = __fadd_rn(p[1][1], -p[2][1]); inv[1] = __fadd_rn(p[2][0], -p[1][0]); inv[2] = mm(p[1][0], p[2][1], p[2][0], p[1][1]);
        inv[3] = __fadd_rn(p[2][1], -p[0][1]); inv[4] = __fadd_rn(p[0][0], -p[2][0]); inv[5] = mm(p[2][0], p[0][1], p[0][0], p[2][1]);
        inv[6] = __fadd_rn(p[0][1], -p[1][1]); inv[7] = __fadd_rn(p[1][0], -p[0][0]); inv[8] = mm(p[0][0], p[1][1], p[1][0], p[0][1]);
        const float den = __fadd_rn(__fadd_rn(__fmul_rn(p[2][0], __fadd_rn(p[0][1], -p[1][1])),
                                              __fmul_rn(p[0][0], __fadd_rn(p[1][1], -p[2][1]))),
                                    __fmul_rn(p[1][0], __fadd_rn(p[2][1], -p[0][1])));
#pragma unroll
        for (int k = 0; k < 9; k++) inv[k] = __fdiv_rn(inv[k], den);
        const float xmin = fminf(p[0][0], fminf(p[1][0], p[2][0])), xmax = fmaxf(p[0][0], fmaxf(p[1][0], p[2][0]));
        const float ymin = fminf(p[0][1], fminf(p[1][1], p[2][1])), ymax = fmaxf(p[0][1], fmaxf(p[1][1], p[2][1]));
        // one pixel of slack on each side absorbs the rounding of the edge functions
        const float lim = (float)is + 8.f;
        const int x0 = max((int)floorf(fmaxf(xmin, -8.f)) - 1, 0), x1 = min((int)ceilf(fminf(xmax, lim)) + 1, is - 1);
        const int y0 = max((int)floorf(fmaxf(ymin, -8.f)) - 1, 0), y1 = min((int)ceilf(fminf(ymax, lim)) + 1, is - 1);
        if (x0 <= x1 && y0 <= y1) bb = make_int2(x0 | (x1 << 16), y0 | (y1 << 16));
    }
#pragma unroll
    for (int k = 0; k < 9; k++) faces_inv[fo * 9 + k] = inv[k];
    bbox[fo] = bb;
}

// ---------------------------------------------------------------------------------------------
// tile kernel
// ---------------------------------------------------------------------------------------------
struct RasterOut {
    int32_t* face_index_map;   // [N,is,is]
    float* weight_map;         // [N,is,is,3]  screen-space barycentrics (clamped, normalised)
    float* depth_map;          // [N,is,is]
    float* alpha_map;          // [N,is,is] or null
    float* face_inv_map;       // [N,is,is,9] or null
};

struct RasterAttrs {           // fused network.Rasterizer.forward post-processing (all optional outputs)
    const float* v;  const int32_t* f_v;     // [nv,3], [nf,3]
    const float* vt; const int32_t* f_vt;    // [nvt,2], [nf,3]
    const float* vn; const int32_t* f_vn;    // [nvn,3], [nf,3]
    const float* pose_R;                     // [N,3,3]
    const float* pose_t;                     // [N,3]
    float* weight_pc;                        // [N,is,is,3] perspective-correct weights
    float* uv_map;                           // [N,is,is,2]
    float* normal_map;                       // [N,is,is,3]
    float* normal_map_cam;
    float* position_map;
    float* position_map_cam;
    int enabled;
};

// coalesced store of a [TH][TW][C] shared-memory tile into an [N,is,is,C] map (rows of the tile are contiguous in HBM)
template <typename T>
__device__ __forceinline__ void store_tile(T* __restrict__ dst, const T* __restrict__ s, int C, int n, int is, int x0, int y0,
                                           int flip) {
    const int wpx = min(TW, is - x0);
    const int row_elems = wpx * C;
    for (int i = threadIdx.x; i < TH * TW * C; i += NT) {
        const int r = i / (TW * C), c = i - r * (TW * C);
        const int yi = y0 + r;
        if (c < row_elems && yi < is) {
            const int yo = flip ? (is - 1 - yi) : yi;
            dst[(((int64_t)n * is + yo) * is + x0) * C + c] = s[i];
        }
    }
}

__global__ void __launch_bounds__(NT) raster_tile_kernel(const float* __restrict__ faces, const float* __restrict__ faces_inv,
                                                       const int2* __restrict__ bbox, int nf, int is, float near, float far,
                                                       int flip, RasterOut o, RasterAttrs a) {
    __shared__ int s_cnt[CHUNK * (NT / 32)];
    __shared__ int s_ids[CHUNK * NT];
    __shared__ float s_rec[BATCH * 18];
    __shared__ float s_stage[NT * 9];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tx1 = min(x0 + TW - 1, is - 1), ty1 = min(y0 + TH - 1, is - 1);
    const int xi = x0 + (tid % TW), yi = y0 + (tid / TW);
    const bool in_img = xi < is && yi < is;
    // pixel centre in NDC: (2. * i + 1 - is) / is evaluated in double like the reference (:96-97)
    const float yp = (float)((2. * yi + 1 - is) / is);
    const float xp = (float)((2. * xi + 1 - is) / is);
    const float fxi = (float)xi, fyi = (float)yi;

    const float* fbase = faces + (int64_t)n * nf * 9;
    const float* ibase = faces_inv + (int64_t)n * nf * 9;
    const int2* bbase = bbox + (int64_t)n * nf;

    float depth_min = far;
    int face_min = -1;
    float w_min[3] = {0.f, 0.f, 0.f};

    for (int base = 0; base < nf; base += CHUNK * NT) {
        // ---- cull: order-preserving compaction of the faces whose bbox touches this tile ----
        unsigned bal[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            const int f = base + j * NT + tid;
            bool hit = false;
            if (f < nf) {
                const int2 b = __ldg(bbase + f);
                const int bx0 = b.x & 0xffff, bx1 = b.x >> 16, by0 = b.y & 0xffff, by1 = b.y >> 16;
                hit = (bx0 <= bx1) && bx0 <= tx1 && bx1 >= x0 && by0 <= ty1 && by1 >= y0;
            }
            bal[j] = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_cnt[j * (NT / 32) + warp] = __popc(bal[j]);
        }
        __syncthreads();
        int total = 0;
        int my_off[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
#pragma unroll
            for (int w = 0; w < NT / 32; w++) {
                const int c = s_cnt[j * (NT / 32) + w];
                if (w == warp) my_off[j] = total;
                total += c;
            }
        }
        if (total == 0) { __syncthreads(); continue; }
#pragma unroll
        for (int j = 0; j < CHUNK; j++)
            if (bal[j] & (1u << lane)) s_ids[my_off[j] + __popc(bal[j] & ((1u << lane) - 1u))] = base + j * NT + tid;
        __syncthreads();
        // ---- test: survivors in batches, records staged in shared memory ----
        for (int b0 = 0; b0 < total; b0 += BATCH) {
            const int nb = min(BATCH, total - b0);
            for (int i = tid; i < nb * 18; i += NT) {
                const int r = i / 18, k = i - r * 18;
                const int f = s_ids[b0 + r];
                s_rec[i] = (k < 9) ? __ldg(fbase + (int64_t)f * 9 + k) : __ldg(ibase + (int64_t)f * 9 + (k - 9));
            }
            __syncthreads();
            if (in_img) {
                for (int r = 0; r < nb; r++) {
                    const float* f = s_rec + r * 18;
                    const float* fi = f + 9;
                    // inside test (:113-116)
                    if ((__fmul_rn(__fadd_rn(yp, -f[1]), __fadd_rn(f[3], -f[0])) < __fmul_rn(__fadd_rn(xp, -f[0]), __fadd_rn(f[4], -f[1]))) ||
                        (__fmul_rn(__fadd_rn(yp, -f[4]), __fadd_rn(f[6], -f[3])) < __fmul_rn(__fadd_rn(xp, -f[3]), __fadd_rn(f[7], -f[4]))) ||
                        (__fmul_rn(__fadd_rn(yp, -f[7]), __fadd_rn(f[0], -f[6])) < __fmul_rn(__fadd_rn(xp, -f[6]), __fadd_rn(f[1], -f[7]))))
                        continue;
                    float w[3];
                    float wsum = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        w[k] = __fadd_rn(__fadd_rn(__fmul_rn(fi[3 * k], fxi), __fmul_rn(fi[3 * k + 1], fyi)), fi[3 * k + 2]);
                        w[k] = fminf(fmaxf(w[k], 0.f), 1.f);
                        wsum = __fadd_rn(wsum, w[k]);
                    }
#pragma unroll
                    for (int k = 0; k < 3; k++) w[k] = __fdiv_rn(w[k], wsum);
                    const float s = __fadd_rn(__fadd_rn(__fdiv_rn(w[0], f[2]), __fdiv_rn(w[1], f[5])), __fdiv_rn(w[2], f[8]));
                    const float zp = (float)(1. / (double)s);
                    if (zp <= near || far <= zp) continue;
                    if (zp < depth_min) {        // survivors arrive in ascending face order: ties keep the lowest index
                        depth_min = zp;
                        face_min = s_ids[b0 + r];
                        w_min[0] = w[0]; w_min[1] = w[1]; w_min[2] = w[2];
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- G-buffer: staged through shared memory, written as contiguous tile rows ----
    int* s_int = (int*)s_stage;
    s_int[tid] = face_min;
    __syncthreads();
    store_tile(o.face_index_map, s_int, 1, n, is, x0, y0, flip);
    __syncthreads();
    s_stage[tid] = depth_min;
    __syncthreads();
    store_tile(o.depth_map, s_stage, 1, n, is, x0, y0, flip);
    __syncthreads();
    if (o.alpha_map) {
        s_stage[tid] = face_min >= 0 ? 1.f : 0.f;
        __syncthreads();
        store_tile(o.alpha_map, s_stage, 1, n, is, x0, y0, flip);
        __syncthreads();
    }
    s_stage[tid * 3] = w_min[0]; s_stage[tid * 3 + 1] = w_min[1]; s_stage[tid * 3 + 2] = w_min[2];
    __syncthreads();
    store_tile(o.weight_map, s_stage, 3, n, is, x0, y0, flip);
    __syncthreads();
    if (o.face_inv_map) {
#pragma unroll
        for (int k = 0; k < 9; k++) s_stage[tid * 9 + k] = face_min >= 0 ? __ldg(ibase + (int64_t)face_min * 9 + k) : 0.f;
        __syncthreads();
        store_tile(o.face_inv_map, s_stage, 9, n, is, x0, y0, flip);
        __syncthreads();
    }
    if (!a.enabled) return;

    // ---- network.Rasterizer.forward :176-214 ----
    // background pixels index the LAST face with zero weights (python's -1), which yields exact zeros everywhere
    const int fsel = face_min >= 0 ? face_min : nf - 1;
    float pw[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float z = __ldg(fbase + (int64_t)fsel * 9 + 3 * k + 2);
        pw[k] = __fmul_rn(__fmul_rn(__fdiv_rn(1.f, z), w_min[k]), depth_min);
    }
    if (a.weight_pc) {
        s_stage[tid * 3] = pw[0]; s_stage[tid * 3 + 1] = pw[1]; s_stage[tid * 3 + 2] = pw[2];
        __syncthreads();
        store_tile(a.weight_pc, s_stage, 3, n, is, x0, y0, flip);
        __syncthreads();
    }
    auto blend = [&](const float* attr, const int32_t* idx, int A, float* out) {
        const int32_t* id = idx + (int64_t)fsel * 3;
        const float* a0 = attr + (int64_t)__ldg(id + 0) * A;
        const float* a1 = attr + (int64_t)__ldg(id + 1) * A;
        const float* a2 = attr + (int64_t)__ldg(id + 2) * A;
        for (int c = 0; c < A; c++)
            out[c] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(a0 + c), pw[0]), __fmul_rn(__ldg(a1 + c), pw[1])), __fmul_rn(__ldg(a2 + c), pw[2]));
    };
    if (a.uv_map) {
        float uv[2];
        blend(a.vt, a.f_vt, 2, uv);
        s_stage[tid * 2] = __fadd_rn(uv[0], -floorf(uv[0]));
        s_stage[tid * 2 + 1] = __fadd_rn(uv[1], -floorf(uv[1]));
        __syncthreads();
        store_tile(a.uv_map, s_stage, 2, n, is, x0, y0, flip);
        __syncthreads();
    }
    if (a.normal_map || a.normal_map_cam) {
        float nm[3];
        blend(a.vn, a.f_vn, 3, nm);
        normalize3(nm[0], nm[1], nm[2]);
        if (a.normal_map) {
            s_stage[tid * 3] = nm[0]; s_stage[tid * 3 + 1] = nm[1]; s_stage[tid * 3 + 2] = nm[2];
            __syncthreads();
            store_tile(a.normal_map, s_stage, 3, n, is, x0, y0, flip);
            __syncthreads();
        }
        if (a.normal_map_cam) {
            const float* Rn = a.pose_R + n * 9;
            float cx = Rn[0] * nm[0] + Rn[1] * nm[1] + Rn[2] * nm[2];
            float cy = Rn[3] * nm[0] + Rn[4] * nm[1] + Rn[5] * nm[2];
            float cz = Rn[6] * nm[0] + Rn[7] * nm[1] + Rn[8] * nm[2];
            normalize3(cx, cy, cz);
            s_stage[tid * 3] = cx; s_stage[tid * 3 + 1] = cy; s_stage[tid * 3 + 2] = cz;
            __syncthreads();
            store_tile(a.normal_map_cam, s_stage, 3, n, is, x0, y0, flip);
            __syncthreads();
        }
    }
    if (a.position_map || a.position_map_cam) {
        float ps[3];
        blend(a.v, a.f_v, 3, ps);
        if (a.position_map) {
            s_stage[tid * 3] = ps[0]; s_stage[tid * 3 + 1] = ps[1]; s_stage[tid * 3 + 2] = ps[2];
            __syncthreads();
            store_tile(a.position_map, s_stage, 3, n, is, x0, y0, flip);
            __syncthreads();
        }
        if (a.position_map_cam) {
            const float* Rn = a.pose_R + n * 9;
            const float* tn = a.pose_t + n * 3;
            s_stage[tid * 3 + 0] = Rn[0] * ps[0] + Rn[1] * ps[1] + Rn[2] * ps[2] + tn[0];
            s_stage[tid * 3 + 1] = Rn[3] * ps[0] + Rn[4] * ps[1] + Rn[5] * ps[2] + tn[1];
            s_stage[tid * 3 + 2] = Rn[6] * ps[0] + Rn[7] * ps[1] + Rn[8] * ps[2] + tn[2];
            __syncthreads();
            store_tile(a.position_map_cam, s_stage, 3, n, is, x0, y0, flip);
            __syncthreads();
        }
    }
}

}  // namespace

extern "C" int rnr_project_vertices(const float* vertices, int v_batch, int nv, const float* K, const float* R, const float* t,
                                    const float* dist_coeffs, const float* offset, const float* scale, float orig_size, float eps,
                                    float* out_uvz, int N, void* stream) {
    if ((int64_t)N * nv == 0) return 0;
    RNR_REQUIRE(v_batch == 1 || v_batch == N, "rnr_project_vertices: vertex batch must be 1 or N");
    dim3 grid(rnr_cdiv(nv, 256), N);
    project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vertices, v_batch, nv, K, R, t, dist_coeffs, offset, scale, orig_size, eps,
                                                          out_uvz, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_raster_face_setup(const float* uvz, int nv, const int32_t* faces_idx, int f_batch, const float* faces_in, int nf,
                                     int image_size, float* faces_out, float* faces_inv, int32_t* bbox, int N, void* stream) {
    if ((int64_t)N * nf == 0) return 0;
    RNR_REQUIRE(image_size >= 1 && image_size <= 32767, "rasterizer: image_size %d out of range", image_size);
    RNR_REQUIRE(faces_in || (uvz && faces_idx && nv > 0), "rasterizer: need either faces or (uvz, face indices)");
    RNR_REQUIRE(faces_inv && bbox, "rasterizer: faces_inv / bbox outputs are required");
    dim3 grid(rnr_cdiv(nf, 256), N);
    face_setup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(uvz, nv, faces_idx, f_batch, faces_in, nf, image_size, faces_out, faces_inv,
                                                             (int2*)bbox, N);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_raster_tiles(const float* faces, const float* faces_inv, const int32_t* bbox, int nf, int image_size, float near,
                                float far, int flip_y, int32_t* face_index_map, float* weight_map, float* depth_map, float* alpha_map,
                                float* face_inv_map, const rnr_raster_attrs_t* attrs, int N, void* stream) {
    if (N == 0) return 0;
    RNR_REQUIRE(nf > 0, "rasterizer: empty mesh");
    RNR_REQUIRE(face_index_map && weight_map && depth_map, "rasterizer: face_index / weight / depth maps are required");
    RasterOut o = {face_index_map, weight_map, depth_map, alpha_map, face_inv_map};
    RasterAttrs a;
    memset(&a, 0, sizeof(a));
    if (attrs) {
        a.v = attrs->v; a.f_v = attrs->f_v_idx; a.vt = attrs->vt; a.f_vt = attrs->f_vt_idx; a.vn = attrs->vn; a.f_vn = attrs->f_vn_idx;
        a.pose_R = attrs->pose_R; a.pose_t = attrs->pose_t;
        a.weight_pc = attrs->weight_pc; a.uv_map = attrs->uv_map; a.normal_map = attrs->normal_map;
        a.normal_map_cam = attrs->normal_map_cam; a.position_map = attrs->position_map; a.position_map_cam = attrs->position_map_cam;
        a.enabled = 1;
        RNR_REQUIRE(!(a.uv_map) || (a.vt && a.f_vt), "rasterizer: uv_map needs vt / f_vt_idx");
        RNR_REQUIRE(!(a.normal_map || a.normal_map_cam) || (a.vn && a.f_vn), "rasterizer: normal maps need vn / f_vn_idx");
        RNR_REQUIRE(!(a.position_map || a.position_map_cam) || (a.v && a.f_v), "rasterizer: position maps need v / f_v_idx");
        RNR_REQUIRE(!(a.normal_map_cam || a.position_map_cam) || (a.pose_R && a.pose_t), "rasterizer: camera-space maps need the pose");
    }
    dim3 grid(rnr_cdiv(image_size, TW), rnr_cdiv(image_size, TH), N);
    raster_tile_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(faces, faces_inv, (const int2*)bbox, nf, image_size, near, far, flip_y, o, a);
    RNR_LAUNCH_CHECK();
    return 0;
}
