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

// ---------------------------------------------------------------------------------------
// Microbenchmark (profiling aid): dependent-load latency at the step kernel's concurrency.  Every warp
// chases `hops` pointers through a table of 64-byte records (next = rec[0]); returns nothing useful.
// ---------------------------------------------------------------------------------------
__global__ void chase_kernel(const int4 *recs, long long n, int hops, int *sink) {
    long long idx = ((long long)blockIdx.x * 2654435761ll) % n;
    int acc = 0;
    for (int h = 0; h < hops; ++h) {
        const int4 r = recs[idx * 4];          // all lanes, same address: one 16 B broadcast load
        acc += r.y;
        idx = (long long)(unsigned)r.x % n;
    }
    if (threadIdx.x == 0) sink[blockIdx.x] = acc + (int)idx;
}

extern "C" int samd_debug_pointer_chase(const void *recs_dev, int64_t n_records, int n_warps, int hops, int32_t *sink_dev, void *stream) {
    chase_kernel<<<n_warps, 32, 0, (cudaStream_t)stream>>>((const int4 *)recs_dev, (long long)n_records, hops, sink_dev);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
