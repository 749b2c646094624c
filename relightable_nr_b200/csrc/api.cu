// Library info + error plumbing of librnr_b200.so.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void rnr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* rnr_version(void) { return "rnr_b200 0.1 (sm_100a)"; }
extern "C" const char* rnr_last_error(void) { return g_err; }

extern "C" int rnr_device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return n;
}
