// Ray sampling and ray rendering.
//   network.RaySampler.forward  (network.py:445-472; camera.get_reflect_dir camera.py:35-45;
//                                render.spherical_mapping_batch render.py:105-111)
//   network.RayRenderer.forward (network.py:481-527) forward + backward
// HBM-bound per-pixel kernels; module-level tensor layouts are the reference's
// (rays_* : [N,H,W,{3,2},R], rays_lt / rays_color : [N,R,3,H,W], images NCHW).
#include "pixel.cuh"

#define RNR_MAX_RAYS 32

namespace {

// ---------------------------------------------------------------------------------------------
// RaySampler
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ray_sampler_kernel(const float* __restrict__ tbn, const float* __restrict__ vdt,
                                                        const float* __restrict__ alpha, const float* __restrict__ pivots /*[3,R]*/,
                                                        int R, int reflect, float* __restrict__ rays_dir,
                                                        float* __restrict__ rays_uv, float* __restrict__ rays_dir_tan, int64_t P) {
    extern __shared__ float sm[];              // [128][3R] dir, [128][2R] uv, [128][3R] tan
    __shared__ float s_piv[3 * RNR_MAX_RAYS];
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * R; i += 128) s_piv[i] = pivots[i];
    __syncthreads();
    float* s_dir = sm;
    float* s_uv = sm + 128 * 3 * R;
    float* s_tan = s_uv + 128 * 2 * R;
    const int64_t p0 = (int64_t)blockIdx.x * 128;
    const int64_t pix = p0 + tid;
    if (pix < P) {
        float T[9];
#pragma unroll
        for (int i = 0; i < 9; i++) T[i] = tbn[pix * 9 + i];
        const float a = alpha[pix];
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (reflect) { vx = vdt[pix * 3 + 0]; vy = vdt[pix * 3 + 1]; vz = vdt[pix * 3 + 2]; }
        for (int r = 0; r < R; r++) {
            const float px = s_piv[r], py = s_piv[R + r], pz = s_piv[2 * R + r];
            float tx, ty, tz;
            if (reflect) {
                const float d = (px * vx + py * vy + pz * vz) * 2.0f;
                tx = d * px - vx; ty = d * py - vy; tz = d * pz - vz;
                normalize3(tx, ty, tz);
                tx *= a; ty *= a; tz *= a;
                s_tan[tid * 3 * R + 0 * R + r] = tx;
                s_tan[tid * 3 * R + 1 * R + r] = ty;
                s_tan[tid * 3 * R + 2 * R + r] = tz;
            } else {
                tx = px; ty = py; tz = pz;
            }
            float wx = T[0] * tx + T[1] * ty + T[2] * tz;
            float wy = T[3] * tx + T[4] * ty + T[5] * tz;
            float wz = T[6] * tx + T[7] * ty + T[8] * tz;
            normalize3(wx, wy, wz);
            s_dir[tid * 3 * R + 0 * R + r] = wx;
            s_dir[tid * 3 * R + 1 * R + r] = wy;
            s_dir[tid * 3 * R + 2 * R + r] = wz;
            float u, v;
            spherical_uv(wx, wy, wz, u, v);
            const float bg = (a == 0.f) ? 1.f : 0.f;
            s_uv[tid * 2 * R + 0 * R + r] = u * a - bg;
            s_uv[tid * 2 * R + 1 * R + r] = v * a - bg;
        }
    }
    __syncthreads();
    const int64_t np = (P - p0) < 128 ? (P - p0) : 128;
    for (int64_t i = tid; i < np * 3 * R; i += 128) rays_dir[p0 * 3 * R + i] = s_dir[i];
    for (int64_t i = tid; i < np * 2 * R; i += 128) rays_uv[p0 * 2 * R + i] = s_uv[i];
    if (reflect && rays_dir_tan)
        for (int64_t i = tid; i < np * 3 * R; i += 128) rays_dir_tan[p0 * 3 * R + i] = s_tan[i];
}

// ---------------------------------------------------------------------------------------------
// RayRenderer
// ---------------------------------------------------------------------------------------------
struct RRParams {
    const float* alb_s;      // [N,3,H,W]
    const float* alb_d;      // [N,3,H,W] or null
    const float* rays_uv;    // [N,H,W,2,R]
    const float* rays_lt;    // [N,R,3,H,W]
    const float* lp;         // [Nl,Hl,Wl,3]
    int Nl, Hl, Wl;
    int R, Rs, Rd;           // total, specular, diffuse
    int no_albedo, separate;
    int64_t HW;
    int N;
};

__global__ void __launch_bounds__(128) ray_render_fwd_kernel(const RRParams q, float* __restrict__ out, float* __restrict__ out_s,
                                                           float* __restrict__ out_d, float* __restrict__ ltt_s,
                                                           float* __restrict__ ltt_d, float* __restrict__ rays_color) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= q.HW * q.N) return;
    const int n = (int)(pix / q.HW);
    const int64_t p = pix % q.HW;
    const float* L = q.lp + (q.Nl == 1 ? 0 : (int64_t)n * q.Hl * q.Wl * 3);
    float ss[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
    for (int r = 0; r < q.R; r++) {
        const float u = q.rays_uv[pix * 2 * q.R + r], v = q.rays_uv[pix * 2 * q.R + q.R + r];
        const float x = fminf(u * (float)q.Wl, (float)(q.Wl - 1));
        const float y = fminf(v * (float)q.Hl, (float)(q.Hl - 1));
        const Bilin b = bilinear_setup(x, y, q.Wl, q.Hl);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float col = L[b.i00 * 3 + c] * b.w00 + L[b.i10 * 3 + c] * b.w10 + L[b.i01 * 3 + c] * b.w01 + L[b.i11 * 3 + c] * b.w11;
            const int64_t o = (((int64_t)n * q.R + r) * 3 + c) * q.HW + p;
            rays_color[o] = col;
            const float t = q.rays_lt[o] * col;
            if (r < q.Rs) ss[c] += t; else sd[c] += t;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int64_t o = ((int64_t)n * 3 + c) * q.HW + p;
        const float ls = ss[c] / (float)q.Rs;
        const float as = q.alb_s[o];
        const float os = q.no_albedo ? ls : as * ls;
        float ld = 0.f, od = 0.f;
        if (q.Rd > 0) {
            ld = sd[c] / (float)q.Rd;
            if (q.no_albedo) od = ld;
            else od = (q.separate ? q.alb_d[o] : as) * ld;
        }
        ltt_s[o] = ls; ltt_d[o] = ld; out_s[o] = os; out_d[o] = od; out[o] = os + od;
    }
}

__global__ void __launch_bounds__(128) ray_render_bwd_kernel(const RRParams q, const float* __restrict__ g_out,
                                                           const float* __restrict__ g_os, const float* __restrict__ g_od,
                                                           const float* __restrict__ g_ls, const float* __restrict__ g_ld,
                                                           const float* __restrict__ ltt_s, const float* __restrict__ ltt_d,
                                                           float* __restrict__ g_alb_s, float* __restrict__ g_alb_d,
                                                           float* __restrict__ g_lt, float* __restrict__ g_lp) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= q.HW * q.N) return;
    const int n = (int)(pix / q.HW);
    const int64_t p = pix % q.HW;
    float Gls[3], Gld[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int64_t o = ((int64_t)n * 3 + c) * q.HW + p;
        const float go = g_out ? g_out[o] : 0.f;
        const float Gos = go + (g_os ? g_os[o] : 0.f);
        const float God = (q.Rd > 0) ? go + (g_od ? g_od[o] : 0.f) : 0.f;
        const float as = q.alb_s[o];
        float gas = 0.f, gad = 0.f;
        Gls[c] = (g_ls ? g_ls[o] : 0.f);
        Gld[c] = (g_ld && q.Rd > 0 ? g_ld[o] : 0.f);
        if (q.no_albedo) {
            Gls[c] += Gos; Gld[c] += God;
        } else {
            Gls[c] += Gos * as;
            gas += Gos * ltt_s[o];
            if (q.Rd > 0) {
                if (q.separate) { Gld[c] += God * q.alb_d[o]; gad += God * ltt_d[o]; }
                else { Gld[c] += God * as; gas += God * ltt_d[o]; }
            }
        }
        if (g_alb_s) g_alb_s[o] = gas;
        if (g_alb_d) g_alb_d[o] = gad;
        Gls[c] /= (float)q.Rs;
        if (q.Rd > 0) Gld[c] /= (float)q.Rd;
    }
    const float* L = q.lp + (q.Nl == 1 ? 0 : (int64_t)n * q.Hl * q.Wl * 3);
    float* GL = g_lp ? g_lp + (q.Nl == 1 ? 0 : (int64_t)n * q.Hl * q.Wl * 3) : nullptr;
    for (int r = 0; r < q.R; r++) {
        const float u = q.rays_uv[pix * 2 * q.R + r], v = q.rays_uv[pix * 2 * q.R + q.R + r];
        const float x = fminf(u * (float)q.Wl, (float)(q.Wl - 1));
        const float y = fminf(v * (float)q.Hl, (float)(q.Hl - 1));
        const Bilin b = bilinear_setup(x, y, q.Wl, q.Hl);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float G = (r < q.Rs) ? Gls[c] : Gld[c];
            const int64_t o = (((int64_t)n * q.R + r) * 3 + c) * q.HW + p;
            if (g_lt) {
                const float col = L[b.i00 * 3 + c] * b.w00 + L[b.i10 * 3 + c] * b.w10 + L[b.i01 * 3 + c] * b.w01 + L[b.i11 * 3 + c] * b.w11;
                g_lt[o] = G * col;
            }
            if (GL) {
                const float gc = G * q.rays_lt[o];
                if (gc != 0.f) {
                    if (b.w00 != 0.f) atomicAdd(GL + b.i00 * 3 + c, gc * b.w00);
                    if (b.w10 != 0.f) atomicAdd(GL + b.i10 * 3 + c, gc * b.w10);
                    if (b.w01 != 0.f) atomicAdd(GL + b.i01 * 3 + c, gc * b.w01);
                    if (b.w11 != 0.f) atomicAdd(GL + b.i11 * 3 + c, gc * b.w11);
                }
            }
        }
    }
}

}  // namespace

extern "C" int rnr_ray_sampler_fwd(const float* tbn, const float* vdt, const float* alpha, const float* pivots, int R,
                                   int reflect, float* rays_dir, float* rays_uv, float* rays_dir_tangent, int64_t P,
                                   void* stream) {
    RNR_REQUIRE(R >= 1 && R <= RNR_MAX_RAYS, "ray sampler: 1..%d rays supported, got %d", RNR_MAX_RAYS, R);
    const size_t smem = (size_t)128 * 8 * R * sizeof(float);
    RNR_ONCE_PER_DEVICE({ RNR_CHECK(cudaFuncSetAttribute(ray_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 8 * RNR_MAX_RAYS * 4)); });
    ray_sampler_kernel<<<rnr_cdiv(P, 128), 128, smem, (cudaStream_t)stream>>>(tbn, vdt, alpha, pivots, R, reflect, rays_dir, rays_uv,
                                                                            rays_dir_tangent, P);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_ray_render_fwd(const float* alb_s, const float* alb_d, const float* rays_uv, const float* rays_lt,
                                  const float* lp, int Nl, int Hl, int Wl, int R, int Rd, int no_albedo, int separate,
                                  float* out, float* out_s, float* out_d, float* ltt_s, float* ltt_d, float* rays_color,
                                  int N, int H, int W, void* stream) {
    RNR_REQUIRE(Nl == 1 || Nl == N, "ray renderer: light probe batch must be 1 or N");
    RNR_REQUIRE(!(separate && !no_albedo && Rd > 0 && !alb_d), "ray renderer: seperate_albedo needs albedo_diffuse");
    RRParams q = {alb_s, alb_d, rays_uv, rays_lt, lp, Nl, Hl, Wl, R, R - Rd, Rd, no_albedo, separate, (int64_t)H * W, N};
    ray_render_fwd_kernel<<<rnr_cdiv((int64_t)N * H * W, 128), 128, 0, (cudaStream_t)stream>>>(q, out, out_s, out_d, ltt_s, ltt_d, rays_color);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_ray_render_bwd(const float* alb_s, const float* alb_d, const float* rays_uv, const float* rays_lt,
                                  const float* lp, int Nl, int Hl, int Wl, int R, int Rd, int no_albedo, int separate,
                                  const float* g_out, const float* g_os, const float* g_od, const float* g_ls, const float* g_ld,
                                  const float* ltt_s, const float* ltt_d, float* g_alb_s, float* g_alb_d, float* g_lt, float* g_lp,
                                  int N, int H, int W, void* stream) {
    RRParams q = {alb_s, alb_d, rays_uv, rays_lt, lp, Nl, Hl, Wl, R, R - Rd, Rd, no_albedo, separate, (int64_t)H * W, N};
    ray_render_bwd_kernel<<<rnr_cdiv((int64_t)N * H * W, 128), 128, 0, (cudaStream_t)stream>>>(q, g_out, g_os, g_od, g_ls, g_ld, ltt_s,
                                                                                            ltt_d, g_alb_s, g_alb_d, g_lt, g_lp);
    RNR_LAUNCH_CHECK();
    return 0;
}
