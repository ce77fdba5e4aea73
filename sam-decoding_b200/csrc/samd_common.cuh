// Shared device/host definitions of the flat suffix-automaton layout (sm_100a).
//
// Layout (identical for the per-request dynamic arenas and the static automaton):
//   state record  int4  {link, length, min_endpos, edge_head}          16 B
//   edge slot     uint4 {state, token, target, next_edge_of_state}     16 B
//   edge table    open addressing, 8-slot (128 B = one cache line) buckets, linear over
//                 buckets; a slot is free when .x == SAMD_EMPTY.  Nothing is ever deleted,
//                 so a probe ends at the first bucket that still has a free slot.
//   text          int32, 1-based, text[0] = -1                         (dyn_sam.py:20)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SAMD_EMPTY 0xFFFFFFFFu
#define SAMD_NIL   0xFFFFFFFFu
#define SAMD_FULL  0xFFFFFFFFu
#define SAMD_BUCKET 8

enum { META_NSTATES = 0, META_LAST = 1, META_N = 2, META_CUR = 3, META_CURLEN = 4, META_NEDGES = 5,
       META_OVERFLOW = 6, META_NCLONES = 7, META_HOPS = 8, META_PROBES = 9, META_WORDS = 16 };

__host__ __device__ __forceinline__ uint32_t samd_hash(uint32_t state, uint32_t tok) {
    uint32_t h = state * 0x9E3779B1u + tok * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

static inline uint64_t samd_next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}
// edge-table capacity rule: >= 4 slots per token (edges <= 3n-4), power of two, >= 64
static inline uint64_t samd_table_slots(uint64_t n_tokens) {
    uint64_t c = samd_next_pow2(4 * (n_tokens + 1));
    return c < 64 ? 64 : c;
}

struct DynArena {
    int4     *states;   // [B][s_cap]
    uint4    *slots;    // [B][h_cap]
    int32_t  *text;     // [B][t_cap]
    int32_t  *meta;     // [B][META_WORDS]
    int32_t   n_requests, max_tokens;
    uint32_t  s_cap, h_cap, t_cap, bmask;
};

struct StaticDev {
    const int4    *states;
    const uint4   *slots;
    const int32_t *text;
    const int32_t *occ;    // cnt_endpos (count flavour) or NULL
    const int2    *topk;   // [n_states][8] (token,target), -1 padded, or NULL
    int64_t  n_states, n_slots, n_tokens;
    uint32_t bmask;
};

struct samd_static_s {
    StaticDev dev;
    // host mirrors kept for export/save (test + persistence paths)
    int4     *h_states;
    uint4    *h_slots;
    int32_t  *h_text;
    int32_t  *h_occ;
    int2     *h_topk;
    int64_t   n_edges, n_clones;
    int       with_counts;
    int       device;
};

struct samd_dyn_s {
    DynArena a;
    int device;
    int64_t bytes;
};

void samd_set_error(const char *fmt, ...);
void samd_count_launch(int n = 1);

#define SAMD_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            samd_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

#define SAMD_REQUIRE(cond, msg)                                       \
    do {                                                              \
        if (!(cond)) {                                                \
            samd_set_error("%s (%s:%d)", msg, __FILE__, __LINE__);    \
            return 2;                                                 \
        }                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------
// Warp-cooperative probe: lanes 0..7 read the eight slots of one bucket (one 128 B line),
// lane 8 fetches the state record of `state` in the same memory round trip.
// ---------------------------------------------------------------------------------------
struct Probe {
    uint32_t slot;     // matching slot, or the first free slot of the probe sequence
    uint32_t target;   // transition target when found
    int4     rec;      // state record of `state` (when requested)
    bool     found;
};

template <bool kRec, bool kReadOnly>
__device__ __forceinline__ Probe warp_probe(const uint4 *slots, uint32_t bmask, const int4 *states, uint32_t state,
                                            uint32_t tok, int lane) {
    Probe r;
    uint32_t b = samd_hash(state, tok) & bmask;
    int4 rec = make_int4(0, 0, 0, 0);
    if (kRec && lane == 8) rec = kReadOnly ? __ldg(states + state) : states[state];
    while (true) {
        uint4 s = make_uint4(0xFFFFFFFEu, 0, 0, 0);
        if (lane < SAMD_BUCKET) {
            const uint4 *ptr = slots + ((size_t)b * SAMD_BUCKET + lane);
            s = kReadOnly ? __ldg(ptr) : *ptr;
        }
        unsigned hit = __ballot_sync(SAMD_FULL, s.x == state && s.y == tok);
        unsigned emp = __ballot_sync(SAMD_FULL, s.x == SAMD_EMPTY);
        if (hit) {
            int l = __ffs(hit) - 1;
            r.found = true;
            r.slot = b * SAMD_BUCKET + l;
            r.target = __shfl_sync(SAMD_FULL, s.z, l);
            break;
        }
        if (emp) {
            int l = __ffs(emp) - 1;
            r.found = false;
            r.slot = b * SAMD_BUCKET + l;
            r.target = 0;
            break;
        }
        b = (b + 1) & bmask;
    }
    if (kRec) {
        r.rec.x = __shfl_sync(SAMD_FULL, rec.x, 8);
        r.rec.y = __shfl_sync(SAMD_FULL, rec.y, 8);
        r.rec.z = __shfl_sync(SAMD_FULL, rec.z, 8);
        r.rec.w = __shfl_sync(SAMD_FULL, rec.w, 8);
    } else {
        r.rec = rec;
    }
    return r;
}

// Prefetch hint for a probe whose key is already known (next token's cursor / tail probes): the
// four sectors of the bucket and the state record are pulled towards the SM while the warp is
// still busy with the current token, turning a DRAM round trip into a cache hit.
__device__ __forceinline__ void prefetch_probe(const uint4 *slots, uint32_t bmask, const int4 *states, uint32_t state,
                                               uint32_t tok, int lane) {
    if (lane < 4) {
        const char *p = reinterpret_cast<const char *>(slots + (size_t)(samd_hash(state, tok) & bmask) * SAMD_BUCKET) + lane * 32;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    } else if (lane == 4 && states) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(states + state));
    }
}

// Longest-suffix-match step (dyn_sam.py:69-78): one memory round trip per suffix-link hop.
template <bool kReadOnly>
__device__ __forceinline__ void warp_transfer(const uint4 *slots, uint32_t bmask, const int4 *states, int &index,
                                              int &length, int tok, int lane, int &hops) {
    bool first = true;
    while (true) {
        Probe pr = warp_probe<true, kReadOnly>(slots, bmask, states, (uint32_t)index, (uint32_t)tok, lane);
        hops++;
        if (!first) length = pr.rec.y;           // length = states[index].length after a link hop
        if (pr.found) {
            index = (int)pr.target;
            length += 1;
            return;
        }
        if (index == 0) {
            length = 0;
            return;
        }
        index = pr.rec.x;
        first = false;
    }
}
