// Library-wide C-ABI plumbing: version, error string, launch counter.
#include "samd_common.cuh"
#include "../../include/samd_b200.h"

#include <atomic>
#include <cstdarg>

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void samd_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void samd_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int samd_abi_version(void) { return SAMD_ABI_VERSION; }
extern "C" const char *samd_last_error(void) { return g_error; }
extern "C" int64_t samd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int samd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
