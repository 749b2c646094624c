#include <stdlib.h>
// Library info + error plumbing of librnr_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[1024] = "";

void rnr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void rnr_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long rnr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* rnr_version(void) { return "rnr_b200 0.1 (sm_100a)"; }
extern "C" const char* rnr_last_error(void) { return g_err; }

extern "C" int rnr_device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return n;
}

// RNR_PDL=0 disables programmatic dependent launch (common.cuh); read once.
int rnr_pdl_enabled(void) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RNR_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
