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

// ---- programmatic dependent launch ------------------------------------------------------------
// The training step is a chain of ~145 short dependent launches; with plain stream order every one of them pays the launch latency
// and its own prologue (barrier / TMEM set-up, constant loads) AFTER its predecessor has drained.  Kernels on that chain are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization (RNR_PDL_LAUNCH) and
//   * call pdl_launch_dependents() first: once every CTA of the grid has started, the NEXT kernel of the stream may be scheduled
//     onto whatever SM resources are free;
//   * call pdl_wait() before they touch anything a predecessor wrote (or overwrite anything it reads): it returns when all
//     prerequisite grids have completed and flushed.  Everything above the wait -- launch latency, set-up, loads of plan
//     constants -- overlaps the predecessor's tail.
// A kernel that takes the attribute MUST execute pdl_wait(); without the attribute both instructions are no-ops.
// RNR_PDL=0 launches everything with plain stream order.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
int rnr_pdl_enabled(void);
template <typename... KArgs, typename... Args>
static inline cudaError_t rnr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rnr_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define RNR_PDL_LAUNCH(kernel, grid, block, smem, stream, ...)                                         \
    do {                                                                                               \
        cudaError_t _le = rnr_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__); \
        if (_le != cudaSuccess) {                                                                      \
            rnr_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_le));  \
            return (int)_le;                                                                           \
        }                                                                                              \
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
