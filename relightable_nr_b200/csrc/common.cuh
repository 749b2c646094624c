// Shared device/host helpers for librnr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/rnr_b200.h"

// ---- error plumbing -------------------------------------------------------------------------
void rnr_set_error(const char* fmt, ...);
void rnr_count_launch(void);          // every kernel launch of the library is counted (rnr_launch_count)

#define RNR_CHECK(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            rnr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

#define RNR_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            rnr_set_error(__VA_ARGS__);                                                   \
            return (int)cudaErrorInvalidValue;                                            \
        }                                                                                 \
    } while (0)

#define RNR_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        rnr_count_launch();                                                               \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            rnr_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

// Runs the statements once per CUDA device (function attributes such as MaxDynamicSharedMemorySize are per device; a
// process-wide flag would leave a second GPU of the same process without the opt-in).  Racing threads may both run it: the
// attribute calls are idempotent.
#define RNR_ONCE_PER_DEVICE(...)                                                          \
    do {                                                                                  \
        static unsigned char _rnr_done[64] = {0};                                         \
        int _rnr_dev = 0;                                                                 \
        RNR_CHECK(cudaGetDevice(&_rnr_dev));                                              \
        if (_rnr_dev < 0 || _rnr_dev >= 64 || !_rnr_done[_rnr_dev]) {                     \
            __VA_ARGS__;                                                                  \
            if (_rnr_dev >= 0 && _rnr_dev < 64) _rnr_done[_rnr_dev] = 1;                  \
        }                                                                                 \
    } while (0)

static inline int rnr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int rnr_dtype_size(int dt) { return dt == RNR_F32 ? 4 : 2; }

// ---- 16-bit <-> float -----------------------------------------------------------------------
__device__ __forceinline__ float ld16(const void* p, int dtype) {
    if (dtype == RNR_F16) return __half2float(*(const __half*)p);
    return __bfloat162float(*(const __nv_bfloat16*)p);
}
__device__ __forceinline__ float cvt16(unsigned short bits, int dtype) {
    if (dtype == RNR_F16) return __half2float(__ushort_as_half(bits));
    return __bfloat162float(__ushort_as_bfloat16(bits));
}
__device__ __forceinline__ unsigned short f2b16(float v, int dtype) {
    if (dtype == RNR_F16) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ void st_out(void* base, int64_t idx, float v, int dtype) {
    if (dtype == RNR_F32) ((float*)base)[idx] = v;
    else ((unsigned short*)base)[idx] = f2b16(v, dtype);
}
__device__ __forceinline__ float ld_any(const void* base, int64_t idx, int dtype) {
    if (dtype == RNR_F32) return ((const float*)base)[idx];
    return cvt16(((const unsigned short*)base)[idx], dtype);
}

// reflect index for ReflectionPad2d(1): padded coordinate p in [0, n+2) -> source in [0,n)
__device__ __forceinline__ int reflect1(int p, int n) {
    int i = p - 1;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// device-side copies of the views (same layout as rnr_view_t)
struct ViewD {
    const void* ptr;
    int32_t dim[4];
    int64_t stride[4];
};
