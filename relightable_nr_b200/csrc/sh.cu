// Spherical-harmonic lighting kernels.
//   sph_harm.evaluate_sh_basis(lmax=2)   sph_harm.py:41-71  (per-pixel, replaces the CPU/pyshtools round trip
//                                        of test_rnr.py:322-329 / precompute.py:239)
//   sph_harm.reconstruct_sh              sph_harm.py:91-102 (basis x coeff inner product; LightingSH.reconstruct_lp
//                                        network.py:622-627) forward + backward
//   sph_harm.fit_sh_coeff                sph_harm.py:74-88  (Monte-Carlo projection 4*pi/N * sum samples*basis)
// The inner products are warp-shuffle reductions (one warp per sample point / per basis row), HBM-bound on the
// [P, num_basis] basis table.
#include "pixel.cuh"

namespace {

// real orthonormal SH, no Condon-Shortley phase, order (l, m=-l..l), m<0 <-> sin(|m| phi)   
__global__ void __launch_bounds__(256) sh_basis_l2_kernel(const float* __restrict__ dirs, float* __restrict__ out, int64_t P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float x = dirs[i * 3 + 0], y = dirs[i * 3 + 1], z = dirs[i * 3 + 2];
    const float r = sqrtf(x * x + y * y + z * z);
    if (r > 0.f) { x /= r; y /= r; z /= r; } else { x = 1.f; y = 0.f; z = 0.f; }   // atan2(0,0)=0 -> azimuth 0, colatitude 90
    float* o = out + i * 9;
    o[0] = 0.28209479177387814f;
    o[1] = 0.4886025119029199f * y;
    o[2] = 0.4886025119029199f * z;
    o[3] = 0.4886025119029199f * x;
    o[4] = 1.0925484305920792f * x * y;
    o[5] = 1.0925484305920792f * y * z;
    o[6] = 0.31539156525252005f * (3.f * z * z - 1.f);
    o[7] = 1.0925484305920792f * x * z;
    o[8] = 0.5462742152960396f * (x * x - y * y);
}


// General-degree basis table (LightingSH.__init__ network.py:557,581; LightingLP.fit_sh :696 call
// sph_harm.evaluate_sh_basis(lmax=10) once at set-up).  fp64 throughout, one thread per direction; the
// associated Legendre functions are built per order m by the stable upward recurrence in l, so no
// per-thread table is needed.  out [P, (lmax+1)^2] double, same (l, m=-l..l) order as the l2 kernel.
__global__ void __launch_bounds__(128) sh_basis_general_kernel(const float* __restrict__ dirs, double* __restrict__ out,
                                                             int64_t P, int lmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double x = dirs[i * 3 + 0], y = dirs[i * 3 + 1], z = dirs[i * 3 + 2];
    const double rxy = sqrt(x * x + y * y);
    const double r = sqrt(rxy * rxy + z * z);
    const double ct = r > 0.0 ? z / r : 0.0, st = r > 0.0 ? rxy / r : 1.0;
    const double phi = atan2(y, x);
    const int nb = (lmax + 1) * (lmax + 1);
    double* o = out + i * nb;
    double pmm = 1.0;                                  // P_m^m without Condon-Shortley phase
    for (int m = 0; m <= lmax; m++) {
        if (m > 0) pmm *= (2 * m - 1) * st;
        const double cm = cos(m * phi), sm = sin(m * phi);
        double p2 = 0.0, p1 = pmm;                     // P_{l-2}^m, P_{l-1}^m
        for (int l = m; l <= lmax; l++) {
            double pl;
            if (l == m) pl = pmm;
            else if (l == m + 1) pl = (2 * m + 1) * ct * pmm;
            else pl = ((2 * l - 1) * ct * p1 - (l + m - 1) * p2) / (l - m);
            if (l > m) { p2 = p1; p1 = pl; }
            const double nrm = sqrt((m == 0 ? 1.0 : 2.0) * (2 * l + 1) / (4.0 * 3.14159265358979323846) *
                                    exp(lgamma((double)(l - m + 1)) - lgamma((double)(l + m + 1))));
            o[l * l + l + m] = nrm * pl * cm;
            if (m > 0) o[l * l + l - m] = nrm * pl * sm;
        }
    }
}

// out[l, p, c] = sum_b basis[p, b] * coeff[l, b, c]        one warp per (l, p)
__global__ void __launch_bounds__(256) sh_reconstruct_kernel(const float* __restrict__ basis, const float* __restrict__ coeff,
                                                           float* __restrict__ out, int64_t P, int B, int Cc, int Lc, int ldo) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= P * Lc) return;
    const int l = (int)(wid / P);
    const int64_t pp = wid % P;
    const float* brow = basis + pp * B;
    const float* cf = coeff + (int64_t)l * B * Cc;
    for (int c0 = 0; c0 < Cc; c0 += 4) {
        float a[4] = {0, 0, 0, 0};
        for (int b = lane; b < B; b += 32) {
            const float bv = brow[b];
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c0 + c < Cc) a[c] += bv * cf[b * Cc + c0 + c];
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c0 + c < Cc) out[((int64_t)l * P + pp) * ldo + c0 + c] = a[c];
        }
    }
}

// res[l, b, c] += scale * sum_p basis[p, b] * v[l, p, c]     (backward of reconstruct, and fit_sh_coeff)
// block = 128 threads (one basis function each, looping if B > 128) x a chunk of points
__global__ void __launch_bounds__(128) sh_project_kernel(const float* __restrict__ basis, float* __restrict__ v,
                                                       float* __restrict__ res, int64_t P, int B, int Cc, int Lc, float scale,
                                                       int chunk, int ldv, int zero_v) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_v[];    // [chunk][Cc]
    const int l = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * chunk;
    const int64_t np = (P - p0) < chunk ? (P - p0) : chunk;
    if (ldv == Cc && !zero_v) {
        for (int64_t i = threadIdx.x; i < np * Cc; i += 128) s_v[i] = v[((int64_t)l * P + p0) * Cc + i];
    } else {
        // pitched rows (ldv >= Cc); zero_v: the accumulator is cleared on the way (every element is read by exactly one block)
        if (ldv == 4 && ((((uintptr_t)v) & 15) == 0)) {
            float4* v4 = (float4*)v + ((int64_t)l * P + p0);
            for (int64_t pp = threadIdx.x; pp < np; pp += 128) {
                const float4 t = v4[pp];
                const float tt[4] = {t.x, t.y, t.z, t.w};
                for (int c = 0; c < Cc; c++) s_v[pp * Cc + c] = tt[c];
                if (zero_v) v4[pp] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int64_t i = threadIdx.x; i < np * ldv; i += 128) {
                const int64_t pp = i / ldv;
                const int c = (int)(i - pp * ldv);
                float* src = v + ((int64_t)l * P + p0) * ldv + i;
                if (c < Cc) s_v[pp * Cc + c] = *src;
                if (zero_v) *src = 0.f;
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += 128) {
        for (int c0 = 0; c0 < Cc; c0 += 4) {
            float a[4] = {0, 0, 0, 0};
            // a pure stream over the basis table (484 B per point at lmax 10): 8 independent loads in flight per thread -- with
            // one it ran at 1/10 of the copy bandwidth (the loop is load-latency bound, not FMA bound)
            int64_t i = 0;
            for (; i + 8 <= np; i += 8) {
                float bv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) bv[u] = __ldcs(basis + (p0 + i + u) * B + b);
#pragma unroll
                for (int u = 0; u < 8; u++) {
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        if (c0 + c < Cc) a[c] += bv[u] * s_v[(i + u) * Cc + c0 + c];
                }
            }
            for (; i < np; i++) {
                const float bv = basis[(p0 + i) * B + b];
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c0 + c < Cc) a[c] += bv * s_v[i * Cc + c0 + c];
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c0 + c < Cc) atomicAdd(res + ((int64_t)l * B + b) * Cc + c0 + c, a[c] * scale);
        }
    }
}

}  // namespace

extern "C" int rnr_sh_basis_l2(const float* dirs, float* out, int64_t P, void* stream) {
    if (P == 0) return 0;
    sh_basis_l2_kernel<<<rnr_cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(dirs, out, P);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_sh_basis(const float* dirs, double* out, int64_t P, int lmax, void* stream) {
    RNR_REQUIRE(lmax >= 0 && lmax <= 32, "rnr_sh_basis: lmax %d out of range [0, 32]", lmax);
    if (P == 0) return 0;
    sh_basis_general_kernel<<<rnr_cdiv(P, 128), 128, 0, (cudaStream_t)stream>>>(dirs, out, P, lmax);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_sh_reconstruct(const float* basis, const float* coeff, float* out, int64_t P, int B, int Cc, int Lc,
                                  void* stream) {
    if (P * Lc == 0) return 0;
    RNR_PDL_LAUNCH(sh_reconstruct_kernel, rnr_cdiv(P * Lc * 32, 256), 256, 0, stream, basis, coeff, out, P, B, Cc, Lc, Cc);
    RNR_LAUNCH_CHECK();
    return 0;
}

// same with an output row pitch ldo >= Cc (columns >= Cc are left untouched): the fused step keeps the envmap as [P, 4] texels
extern "C" int rnr_sh_reconstruct_ld(const float* basis, const float* coeff, float* out, int64_t P, int B, int Cc, int ldo,
                                     void* stream) {
    RNR_REQUIRE(ldo >= Cc, "rnr_sh_reconstruct_ld: ldo %d < Cc %d", ldo, Cc);
    if (P == 0) return 0;
    RNR_PDL_LAUNCH(sh_reconstruct_kernel, rnr_cdiv(P * 32, 256), 256, 0, stream, basis, coeff, out, P, B, Cc, 1, ldo);
    RNR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rnr_sh_project(const float* basis, const float* v, float* res, int64_t P, int B, int Cc, int Lc, float scale,
                              void* stream) {
    if (P * Lc == 0) return 0;
    int chunk = 256;
    RNR_REQUIRE((size_t)chunk * Cc * 4 <= 48 * 1024, "sh_project: too many channels (%d)", Cc);
    dim3 grid(rnr_cdiv(P, chunk), Lc);
    RNR_PDL_LAUNCH(sh_project_kernel, grid, 128, (size_t)chunk * Cc * 4, stream, basis, const_cast<float*>(v), res, P, B, Cc, Lc, scale, chunk, Cc, 0);
    RNR_LAUNCH_CHECK();
    return 0;
}

// res[b, c] += scale * sum_p basis[p, b] * v[p*ldv + c] for c < Cc; with zero_v the (accumulator) rows of v are cleared as they
// are consumed -- the envmap-gradient texels [P, 4] of the fused step need no separate copy / memset
extern "C" int rnr_sh_project_ld(const float* basis, float* v, float* res, int64_t P, int B, int Cc, int ldv, float scale, int zero_v,
                                 void* stream) {
    RNR_REQUIRE(ldv >= Cc, "rnr_sh_project_ld: ldv %d < Cc %d", ldv, Cc);
    if (P == 0) return 0;
    int chunk = 256;
    RNR_REQUIRE((size_t)chunk * Cc * 4 <= 48 * 1024, "sh_project: too many channels (%d)", Cc);
    dim3 grid(rnr_cdiv(P, chunk), 1);
    RNR_PDL_LAUNCH(sh_project_kernel, grid, 128, (size_t)chunk * Cc * 4, stream, basis, v, res, P, B, Cc, 1, scale, chunk, ldv, zero_v);
    RNR_LAUNCH_CHECK();
    return 0;
}
