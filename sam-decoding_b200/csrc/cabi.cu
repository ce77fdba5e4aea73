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

// ---------------------------------------------------------------------------------------
// Microbenchmark (profiling aid): the ceiling of scattered row moves.  n_granules copies of granule_bytes each,
// base + src_off[g] -> base + dst_off[g], flattened into 16-byte units over the whole grid, four independent units in
// flight per lane - nothing else in the kernel.  tools/row_move_ceiling.py sweeps the granule size at a fixed number
// of bytes: what c4's KV compaction (256-byte granules, one per (tensor, head, row)) can reach at best.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) granule_copy_kernel(char *base, const long long *__restrict__ src_off,
                                                           const long long *__restrict__ dst_off, long long n_granules, int cols) {
    const long long total = n_granules * cols;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; u + 3 * stride < total; u += 4 * stride) {
        uint4 v[4];
        long long d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long w = u + i * stride, g = w / cols, c = w - g * cols;
            v[i] = *reinterpret_cast<const uint4 *>(base + src_off[g] + (c << 4));
            d[i] = dst_off[g] + (c << 4);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4 *>(base + d[i]) = v[i];
    }
    for (; u < total; u += stride) {
        const long long g = u / cols, c = u - g * cols;
        *reinterpret_cast<uint4 *>(base + dst_off[g] + (c << 4)) = *reinterpret_cast<const uint4 *>(base + src_off[g] + (c << 4));
    }
}

extern "C" int samd_debug_granule_copy(void *base_dev, const int64_t *src_off_dev, const int64_t *dst_off_dev, int64_t n_granules,
                                       int32_t granule_bytes, int32_t n_blocks, void *stream) {
    SAMD_REQUIRE(base_dev && src_off_dev && dst_off_dev && n_granules > 0 && granule_bytes >= 16 && granule_bytes % 16 == 0 && n_blocks > 0,
                 "samd_debug_granule_copy: bad arguments");
    granule_copy_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>((char *)base_dev, (const long long *)src_off_dev,
                                                                   (const long long *)dst_off_dev, (long long)n_granules, granule_bytes >> 4);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// Profiling aid: where the hardware puts the warps of a launch shaped like the step kernel's.  Every warp records its
// %smid and %warpid (the SM's warp slot; slot mod 4 = scheduler partition) and lingers for `spin_ns` so that the whole grid
// is resident together.  tools/warp_slots.py
// ---------------------------------------------------------------------------------------
__global__ void warp_slots_kernel(int32_t *out, unsigned spin_ns) {
    if ((threadIdx.x & 31) == 0) {
        unsigned smid, warpid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        out[2 * w] = (int)smid;
        out[2 * w + 1] = (int)warpid;
    }
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < spin_ns);
}

extern "C" int samd_debug_warp_slots(int32_t *out_dev, int n_blocks, int threads, int spin_ns, void *stream) {
    SAMD_REQUIRE(out_dev && n_blocks > 0 && threads > 0 && threads % 32 == 0, "samd_debug_warp_slots: bad arguments");
    warp_slots_kernel<<<n_blocks, threads, 0, (cudaStream_t)stream>>>(out_dev, (unsigned)spin_ns);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
