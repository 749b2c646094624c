// The cold entry points of neural_renderer.cuda (SURVEY.md 8b "B2"): everything in the reference's CUDA extension besides
// forward_face_index_map (raster.cu).  None of them is on the relighting path -- the rasterizer's rgb output is discarded
// (network.py:157), it is never differentiated, and textures are never loaded from / saved to OBJ materials by the scripts --
// but they are part of the extension's surface, so a user of neural_renderer finds all seven functions.
//
//   forward_texture_sampling   rasterize_cuda_kernel.cu:172-242   per-pixel trilinear blend of the face's ts^3 rgb texture cube
//   backward_pixel_map         rasterize_cuda_kernel.cu:245-505   silhouette gradient: per face, walk its three edges along x and y
//   backward_textures          rasterize_cuda_kernel.cu:507-541   scatter of grad_rgb into the texture cubes
//   backward_depth_map         rasterize_cuda_kernel.cu:543-591   depth gradient w.r.t. the face's vertices
//   load_textures              load_textures_cuda_kernel.cu:25-121  texture cubes from an image through per-face uv triangles
//   create_texture_image       create_texture_image_cuda_kernel.cu:9-117   texture cubes -> tiled atlas image
//
// Arithmetic follows the reference expression by expression, including where its float templates promote to double through a
// double literal (`* 2. / is`, `max(x, 0.)`, `/ (ts - 1.)`): the discrete decisions (ceil / floor / truncation of edge crossings)
// must come out the same.  Checked on the GPU against the reference's own kernels recompiled for sm_100a (tests/test_b2_gpu.py).
#include "common.cuh"

namespace {

// ---- the 8-corner blend over a [ts][ts][ts][3] cube shared by texture sampling and the atlas writer -------------------------
struct Corner8 {
    int idx[8];
    float w[8];
};

__device__ __forceinline__ Corner8 cube_corners(const float t[3], int ts) {
    Corner8 c;
    int base[3];
    float frac[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { base[k] = (int)t[k]; frac[k] = t[k] - (float)base[k]; }
#pragma unroll
    for (int pn = 0; pn < 8; pn++) {
        float w = 1.f;
        int id[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const bool hi = (pn >> k) & 1;
            w *= hi ? frac[k] : (1.f - frac[k]);
            id[k] = base[k] + (hi ? 1 : 0);
        }
        c.idx[pn] = id[0] * ts * ts + id[1] * ts + id[2];
        c.w[pn] = w;
    }
    return c;
}

__device__ __forceinline__ float clamp_texel(float v, int ts, float eps) {
    v = (float)fmax((double)v, 0.);
    return fminf(v, (float)(ts - 1) - eps);
}

__global__ void __launch_bounds__(256) texture_sampling_kernel(const float* __restrict__ faces, const float* __restrict__ textures,
                                                             const int32_t* __restrict__ face_index_map, const float* __restrict__ weight_map,
                                                             const float* __restrict__ depth_map, float* __restrict__ rgb_map,
                                                             int32_t* __restrict__ sampling_index_map, float* __restrict__ sampling_weight_map,
                                                             int64_t n_pix, int nf, int is, int ts, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int f = face_index_map[i];
    if (f < 0) return;                                     // background pixels keep whatever the caller put there
    const int64_t bn = i / ((int64_t)is * is);
    const float* face = faces + (bn * nf + f) * 9;
    const float* cube = textures + (bn * nf + f) * (int64_t)ts * ts * ts * 3;
    const float depth = depth_map[i];
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = clamp_texel(weight_map[i * 3 + k] * (float)(ts - 1) * (depth / face[3 * k + 2]), ts, eps);
    const Corner8 c = cube_corners(t, ts);
    float px[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int pn = 0; pn < 8; pn++) {
#pragma unroll
        for (int k = 0; k < 3; k++) px[k] += c.w[pn] * cube[c.idx[pn] * 3 + k];
        sampling_index_map[i * 8 + pn] = c.idx[pn];
        sampling_weight_map[i * 8 + pn] = c.w[pn];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) rgb_map[i * 3 + k] = px[k];
}

__global__ void __launch_bounds__(256) textures_bwd_kernel(const int32_t* __restrict__ face_index_map, const float* __restrict__ sampling_weight_map,
                                                         const int32_t* __restrict__ sampling_index_map, const float* __restrict__ grad_rgb_map,
                                                         float* __restrict__ grad_textures, int64_t n_pix, int nf, int is, int ts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int f = face_index_map[i];
    if (f < 0) return;
    const int64_t bn = i / ((int64_t)is * is);
    float* cube = grad_textures + (bn * nf + f) * (int64_t)ts * ts * ts * 3;
    const float g[3] = {grad_rgb_map[i * 3], grad_rgb_map[i * 3 + 1], grad_rgb_map[i * 3 + 2]};
    for (int pn = 0; pn < 8; pn++) {
        const float w = sampling_weight_map[i * 8 + pn];
        float* dst = cube + (int64_t)sampling_index_map[i * 8 + pn] * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) atomicAdd(dst + k, w * g[k]);
    }
}

__global__ void __launch_bounds__(256) depth_bwd_kernel(const float* __restrict__ faces, const float* __restrict__ depth_map,
                                                      const int32_t* __restrict__ face_index_map, const float* __restrict__ face_inv_map,
                                                      const float* __restrict__ weight_map, const float* __restrict__ grad_depth_map,
                                                      float* __restrict__ grad_faces, int64_t n_pix, int nf, int is) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int f = face_index_map[i];
    if (f < 0) return;
    const int64_t bn = i / ((int64_t)is * is);
    const float* face = faces + (bn * nf + f) * 9;
    float* gface = grad_faces + (bn * nf + f) * 9;
    const float* inv = face_inv_map + i * 9;
    const float* w = weight_map + i * 3;
    const float depth = depth_map[i], d2 = depth * depth, gd = grad_depth_map[i];
    // d depth / d z_k
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float z = face[3 * k + 2];
        atomicAdd(gface + 3 * k + 2, gd * w[k] * d2 / (z * z));
    }
    // d depth / d (x_k, y_k): columns of the inverse face matrix weighted by 1 / z
    float col[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 3; l++) col[k] += -inv[3 * l + k] / face[3 * l + 2];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 2; l++) atomicAdd(gface + 3 * k + l, -gd * col[l] * w[k] * d2 * (float)is / 2);
}

// ---- silhouette gradient -------------------------------------------------------------------------------------------------------
// One thread per (batch, face).  For each of the face's edges, once along x (axis 0: the edge is sampled at integer columns d0 and
// crosses them at row d1_cross) and once along y (axis 1: roles swapped), the pixel just inside the edge and the pixel just
// outside it are compared with the pixels the edge would uncover / cover if it moved: a positive (colour difference x incoming
// gradient) pulls the edge's two end points along the other axis, scaled by the inverse distance to the crossing.
struct EdgeWalk {
    const int32_t* fim;
    const float* rgb;
    const float* alpha;
    const float* grgb;
    const float* galpha;
    int64_t img0;          // bn * is * is
    int is, axis, use_rgb, use_alpha;
    float p[3][2];         // the edge's end points p[0], p[1] and the opposite vertex p[2], (d0, d1) order for this axis
    float eps;

    __device__ __forceinline__ int64_t pix(int d0, int d1) const {
        return axis == 0 ? img0 + (int64_t)d1 * is + d0 : img0 + (int64_t)d0 * is + d1;
    }
    // sum over channels of (value at q - reference value) * incoming gradient at q
    __device__ __forceinline__ float pull(int64_t q, float ref_alpha, const float* ref_rgb) const {
        float d = 0.f;
        if (use_alpha) d += (alpha[q] - ref_alpha) * galpha[q];
        if (use_rgb) {
#pragma unroll
            for (int k = 0; k < 3; k++) d += (rgb[q * 3 + k] - ref_rgb[k]) * grgb[q * 3 + k];
        }
        return d;
    }
    // distribute `d` to the two end points of the edge (component 1 - axis), gradient entries ga (p[0]) and gb (p[1])
    __device__ __forceinline__ void push(float d, int d0, int d1, float d1_cross, float& ga, float& gb) const {
        const float span = p[1][0] - p[0][0];
        if (p[1][0] != (float)d0) {
            float dist = (float)((double)(span / (p[1][0] - (float)d0) * ((float)d1 - d1_cross)) * 2. / is);
            dist = (0.f < dist) ? dist + eps : dist - eps;
            ga -= d / dist;
        }
        if (p[0][0] != (float)d0) {
            float dist = (float)((double)(span / ((float)d0 - p[0][0]) * ((float)d1 - d1_cross)) * 2. / is);
            dist = (0.f < dist) ? dist + eps : dist - eps;
            gb -= d / dist;
        }
    }
};

__global__ void __launch_bounds__(128) pixel_map_bwd_kernel(const float* __restrict__ faces, const int32_t* __restrict__ face_index_map,
                                                          const float* __restrict__ rgb_map, const float* __restrict__ alpha_map,
                                                          const float* __restrict__ grad_rgb_map, const float* __restrict__ grad_alpha_map,
                                                          float* __restrict__ grad_faces, int64_t n_faces_total, int nf, int is, float eps,
                                                          int return_rgb, int return_alpha) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces_total) return;
    const int64_t bn = i / nf;
    const int fn = (int)(i % nf);
    float face[9];
#pragma unroll
    for (int k = 0; k < 9; k++) face[k] = faces[i * 9 + k];
    // back-facing triangles receive no gradient (their entry of grad_faces is left as the caller initialised it)
    if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0])) return;
    float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    EdgeWalk E;
    E.fim = face_index_map; E.rgb = rgb_map; E.alpha = alpha_map; E.grgb = grad_rgb_map; E.galpha = grad_alpha_map;
    E.img0 = bn * (int64_t)is * is; E.is = is; E.use_rgb = return_rgb; E.use_alpha = return_alpha; E.eps = eps;
    for (int e = 0; e < 3; e++) {
        const int v[3] = {e, (e + 1) % 3, (e + 2) % 3};
        float px[3][2];                                    // pixel coordinates of the three vertices, (x, y)
#pragma unroll
        for (int n = 0; n < 3; n++)
#pragma unroll
            for (int d = 0; d < 2; d++) px[n][d] = (float)(0.5 * (double)(face[3 * v[n] + d] * (float)is + (float)is - 1.f));
        for (int axis = 0; axis < 2; axis++) {
            E.axis = axis;
#pragma unroll
            for (int n = 0; n < 3; n++) { E.p[n][0] = px[n][axis]; E.p[n][1] = px[n][1 - axis]; }
            const bool rising = E.p[0][0] < E.p[1][0];
            const int dir = (axis == 0) ? (rising ? -1 : 1) : (rising ? 1 : -1);       // from the inside pixel towards the outside
            const int d0_lo = (int)fmax(ceil((double)fminf(E.p[0][0], E.p[1][0])), 0.);
            const int d0_hi = (int)fmin((double)fmaxf(E.p[0][0], E.p[1][0]), is - 1.);
            float& ga = g[v[0] * 3 + (1 - axis)];
            float& gb = g[v[1] * 3 + (1 - axis)];
            for (int d0 = d0_lo; d0 <= d0_hi; d0++) {
                const float d1_cross = (E.p[1][1] - E.p[0][1]) / (E.p[1][0] - E.p[0][0]) * ((float)d0 - E.p[0][0]) + E.p[0][1];
                const int d1_in = (0 < dir) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
                const int d1_out = d1_in + dir;
                if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                const int64_t q_in = E.pix(d0, d1_in), q_out = E.pix(d0, d1_out);
                float a_in = 0.f, a_out = 0.f, c_in[3] = {0.f, 0.f, 0.f}, c_out[3] = {0.f, 0.f, 0.f};
                if (return_alpha) { a_in = alpha_map[q_in]; a_out = alpha_map[q_out]; }
                if (return_rgb) {
#pragma unroll
                    for (int k = 0; k < 3; k++) { c_in[k] = rgb_map[q_in * 3 + k]; c_out[k] = rgb_map[q_out * 3 + k]; }
                }
                // outward: every pixel from the outside neighbour to the image border, if the inside pixel shows this face
                if (face_index_map[q_in] == fn) {
                    const int lim = (0 < dir) ? is - 1 : 0;
                    const int lo = max(min(d1_out, lim), 0), hi = min(max(d1_out, lim), is - 1);
                    for (int d1 = lo; d1 <= hi; d1++) {
                        const float d = E.pull(E.pix(d0, d1), a_in, c_in);
                        if (d <= 0.f) continue;
                        E.push(d, d0, d1, d1_cross, ga, gb);
                    }
                }
                // inward: the pixels of this face between the inside neighbour and where the column leaves the triangle again
                {
                    float far_cross;
                    if (((float)d0 - E.p[0][0]) * ((float)d0 - E.p[2][0]) < 0.f)
                        far_cross = (E.p[2][1] - E.p[0][1]) / (E.p[2][0] - E.p[0][0]) * ((float)d0 - E.p[0][0]) + E.p[0][1];
                    else
                        far_cross = (E.p[1][1] - E.p[2][1]) / (E.p[1][0] - E.p[2][0]) * ((float)d0 - E.p[2][0]) + E.p[2][1];
                    const int lim = (0 < dir) ? (int)ceilf(far_cross) : (int)floorf(far_cross);
                    const int lo = max(min(d1_in, lim), 0), hi = min(max(d1_in, lim), is - 1);
                    for (int d1 = lo; d1 <= hi; d1++) {
                        const int64_t q = E.pix(d0, d1);
                        if (face_index_map[q] != fn) continue;
                        const float d = E.pull(q, a_out, c_out);
                        if (d <= 0.f) continue;
                        E.push(d, d0, d1, d1_cross, ga, gb);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) grad_faces[i * 9 + k] = g[k];
}

// ---- texture cubes <-> images -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap_mod(float x, float y) { return x > 0.f ? fmodf(x, y) : y + fmodf(x, y); }

__global__ void __launch_bounds__(256) load_textures_kernel(const float* __restrict__ image, const int32_t* __restrict__ is_update,
                                                          float* __restrict__ faces, float* __restrict__ textures, int64_t n_texels, int ts,
                                                          int H, int W, int wrapping, int bilinear) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_texels) return;
    const int64_t cube = (int64_t)ts * ts * ts;
    const int64_t fn = i / cube;
    if (is_update[fn] == 0) return;
    // barycentric position of this texel inside the face
    float b0 = (float)((double)((i / (ts * ts)) % ts) / (ts - 1.));
    float b1 = (float)((double)((i / ts) % ts) / (ts - 1.));
    float b2 = (float)((double)(i % ts) / (ts - 1.));
    if (0.f < b0 + b1 + b2) {
        const float s = b0 + b1 + b2;
        b0 /= s; b1 /= s; b2 /= s;
    }
    // wrap the face's six uv coordinates (values from the ORIGINAL coordinates; texel 0 of the cube writes them back -- the
    // reference lets every thread rewrite them in place, with the same result except for coordinates that are exact integers)
    float uv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float c = faces[fn * 6 + k];
        if (wrapping == 0) c = wrap_mod(c, 1.f);
        else if (wrapping == 1) c = (wrap_mod(c, 2.f) < 1.f) ? wrap_mod(c, 1.f) : 1.f - wrap_mod(c, 1.f);
        else if (wrapping == 2) c = fmaxf(fminf(c, 1.f), 0.f);
        uv[k] = c;
    }
    float* tx = textures + i * 3;
    if (wrapping == 3) { tx[0] = tx[1] = tx[2] = 0.f; return; }     // CLAMP_TO_BORDER: the reference writes zeros
    const float pos_x = (uv[0] * b0 + uv[2] * b1 + uv[4] * b2) * (float)(W - 1);
    const float pos_y = (uv[1] * b0 + uv[3] * b1 + uv[5] * b2) * (float)(H - 1);
    if (bilinear) {
        const int x0 = (int)pos_x, y0 = (int)pos_y;
        const int x1 = min(x0 + 1, W - 1), y1 = min((int)(pos_y + 1.f), H - 1);
        const float wx1 = pos_x - (float)x0, wx0 = 1.f - wx1, wy1 = pos_y - (float)y0, wy0 = 1.f - wy1;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float c = 0.f;
            c += image[((int64_t)y0 * W + x0) * 3 + k] * (wx0 * wy0);
            c += image[((int64_t)y1 * W + x0) * 3 + k] * (wx0 * wy1);
            c += image[((int64_t)y0 * W + x1) * 3 + k] * (wx1 * wy0);
            c += image[((int64_t)y1 * W + x1) * 3 + k] * (wx1 * wy1);
            tx[k] = c;
        }
    } else {
        const int xi = (int)round((double)pos_x), yi = (int)round((double)pos_y);
#pragma unroll
        for (int k = 0; k < 3; k++) tx[k] = image[((int64_t)yi * W + xi) * 3 + k];
    }
}

__global__ void __launch_bounds__(256) wrap_face_uv_kernel(const int32_t* __restrict__ is_update, float* __restrict__ faces, int64_t nf, int wrapping) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf * 6 || wrapping > 2 || is_update[i / 6] == 0) return;
    float c = faces[i];
    if (wrapping == 0) c = wrap_mod(c, 1.f);
    else if (wrapping == 1) c = (wrap_mod(c, 2.f) < 1.f) ? wrap_mod(c, 1.f) : 1.f - wrap_mod(c, 1.f);
    else c = fmaxf(fminf(c, 1.f), 0.f);
    faces[i] = c;
}

__global__ void __launch_bounds__(128) texture_image_kernel(const float* __restrict__ vertices_all, const float* __restrict__ textures,
                                                          float* __restrict__ image, int64_t n_pix, int num_faces, int tsi, int tso,
                                                          int tile_width, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int Wimg = tile_width * tso;
    const int x = (int)(i % Wimg), y = (int)(i / Wimg);
    const int fn = x / tso + (y / tso) * tile_width;
    if (fn >= num_faces) return;      // tiles past the last face stay as the caller initialised them (the reference reads out of bounds there)
    const float* cube = textures + (int64_t)fn * tsi * tsi * tsi * 3;
    const float* vtx = vertices_all + (int64_t)fn * 6;
    const float p0x = vtx[0], p0y = vtx[1], p1x = vtx[2], p1y = vtx[3], p2x = vtx[4], p2y = vtx[5];
    float inv[9] = {p1y - p2y, p2x - p1x, p1x * p2y - p2x * p1y,
                    p2y - p0y, p0x - p2x, p2x * p0y - p0x * p2y,
                    p0y - p1y, p1x - p0x, p0x * p1y - p1x * p0y};
    const float den = p2x * (p0y - p1y) + p0x * (p1y - p2y) + p1x * (p2y - p0y);
#pragma unroll
    for (int k = 0; k < 9; k++) inv[k] /= den;
    float w[3], wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) { w[k] = inv[3 * k] * (float)x + inv[3 * k + 1] * (float)y + inv[3 * k + 2]; wsum += w[k]; }
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = clamp_texel(w[k] / (wsum + eps) * (float)(tsi - 1), tsi, eps);
    const Corner8 c = cube_corners(t, tsi);
    float px[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int pn = 0; pn < 8; pn++)
#pragma unroll
        for (int k = 0; k < 3; k++) px[k] += c.w[pn] * cube[c.idx[pn] * 3 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) image[i * 3 + k] = px[k];
}

// second pass of the atlas writer: the first pixel right of each tile's diagonal copies its left neighbour (seam padding)
__global__ void __launch_bounds__(128) texture_image_seam_kernel(float* __restrict__ image, int64_t n_pix, int tso, int tile_width) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int Wimg = tile_width * tso;
    const int x = (int)(i % Wimg), y = (int)(i / Wimg);
    if ((y % tso + 1) == (x % tso)) {
#pragma unroll
        for (int k = 0; k < 3; k++) image[i * 3 + k] = image[((int64_t)y * Wimg + (x - 1)) * 3 + k];
    }
}

}  // namespace

extern "C" int rnr_nr_forward_texture_sampling(const float* faces, const float* textures, const int32_t* face_index_map, const float* weight_map,
                                               const float* depth_map, float* rgb_map, int32_t* sampling_index_map, float* sampling_weight_map,
                                               int batch, int num_faces, int image_size, int texture_size, float eps, void* stream) {
    RNR_REQUIRE(texture_size >= 2, "forward_texture_sampling: texture_size must be >= 2");
    const int64_t n = (int64_t)batch * image_size * image_size;
    if (n == 0) return 0;
    texture_sampling_kernel<<<rnr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(faces, textures, face_index_map, weight_map, depth_map, rgb_map,
                                                                               sampling_index_map, sampling_weight_map, n, num_faces, image_size,
                                                                               texture_size, eps);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_nr_backward_pixel_map(const float* faces, const int32_t* face_index_map, const float* rgb_map, const float* alpha_map,
                                         const float* grad_rgb_map, const float* grad_alpha_map, float* grad_faces, int batch, int num_faces,
                                         int image_size, float eps, int return_rgb, int return_alpha, void* stream) {
    const int64_t n = (int64_t)batch * num_faces;
    if (n == 0) return 0;
    pixel_map_bwd_kernel<<<rnr_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(faces, face_index_map, rgb_map, alpha_map, grad_rgb_map, grad_alpha_map,
                                                                            grad_faces, n, num_faces, image_size, eps, return_rgb, return_alpha);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_nr_backward_textures(const int32_t* face_index_map, const float* sampling_weight_map, const int32_t* sampling_index_map,
                                        const float* grad_rgb_map, float* grad_textures, int batch, int num_faces, int image_size,
                                        int texture_size, void* stream) {
    const int64_t n = (int64_t)batch * image_size * image_size;
    if (n == 0) return 0;
    textures_bwd_kernel<<<rnr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(face_index_map, sampling_weight_map, sampling_index_map, grad_rgb_map,
                                                                           grad_textures, n, num_faces, image_size, texture_size);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_nr_backward_depth_map(const float* faces, const float* depth_map, const int32_t* face_index_map, const float* face_inv_map,
                                         const float* weight_map, const float* grad_depth_map, float* grad_faces, int batch, int num_faces,
                                         int image_size, void* stream) {
    const int64_t n = (int64_t)batch * image_size * image_size;
    if (n == 0) return 0;
    depth_bwd_kernel<<<rnr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(faces, depth_map, face_index_map, face_inv_map, weight_map,
                                                                        grad_depth_map, grad_faces, n, num_faces, image_size);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_nr_load_textures(const float* image, float* faces, float* textures, const int32_t* is_update, int num_faces, int texture_size,
                                    int image_height, int image_width, int texture_wrapping, int use_bilinear, void* stream) {
    RNR_REQUIRE(texture_size >= 2, "load_textures: texture_size must be >= 2");
    RNR_REQUIRE(texture_wrapping >= 0 && texture_wrapping <= 3, "load_textures: unknown texture_wrapping %d", texture_wrapping);
    const int64_t n = (int64_t)num_faces * texture_size * texture_size * texture_size;
    if (n == 0) return 0;
    load_textures_kernel<<<rnr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(image, is_update, faces, textures, n, texture_size, image_height,
                                                                            image_width, texture_wrapping, use_bilinear);
    RNR_LAUNCH_CHECK();
    // the reference leaves the wrapped uv coordinates in `faces`: same here, after every texel has read the original values
    wrap_face_uv_kernel<<<rnr_cdiv((int64_t)num_faces * 6, 256), 256, 0, (cudaStream_t)stream>>>(is_update, faces, num_faces, texture_wrapping);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_nr_create_texture_image(const float* vertices_all, const float* textures, float* image, int64_t image_numel, int num_faces,
                                           int texture_size_in, int texture_size_out, int tile_width, float eps, void* stream) {
    RNR_REQUIRE(texture_size_in >= 2 && texture_size_out >= 1 && tile_width >= 1, "create_texture_image: bad sizes");
    const int64_t n = image_numel / 3;
    if (n == 0) return 0;
    texture_image_kernel<<<rnr_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(vertices_all, textures, image, n, num_faces, texture_size_in,
                                                                            texture_size_out, tile_width, eps);
    RNR_LAUNCH_CHECK();
    texture_image_seam_kernel<<<rnr_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(image, n, texture_size_out, tile_width);
    RNR_LAUNCH_CHECK();
    return 0;
}
