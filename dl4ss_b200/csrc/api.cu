// Error state, launch counter and device queries shared by every entry point of the C ABI.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace dl4ss {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

}  // namespace dl4ss

extern "C" int dl4ss_version(void) { return 100; }   // 0.1.0
extern "C" const char *dl4ss_last_error(void) { return dl4ss::g_err; }
extern "C" uint64_t dl4ss_launch_count(void) { return dl4ss::g_launches.load(std::memory_order_relaxed); }
