// Document-sharded static SAM without a collective library: the max-reduce of the packed keys
//  done by the look-up kernel itself over NVLink peer memory (sm_100a).
//
// One process per GPU.  Every rank owns one buffer  { keys[2][Q] u64, flags[world], epoch, done counters }  and maps
// every peer's buffer through CUDA IPC.  Per step:
//   look-up kernel   one warp per query walks this shard's automaton, forms the packed key and atomicMax'es it into
//                    EVERY rank's keys[parity][q] (a remote atomic per peer - 8 bytes, performed at the owner's L2);
//                    the last block to finish fences and writes flags[my rank] = epoch on every rank.
//   draft kernel     waits until all flags on THIS rank carry the epoch (every shard's keys have landed), reads
//                    keys[parity][q], clears it for reuse two steps later, and reads the draft from the replicated
//                    corpus.  Its last block publishes the epoch.
// No launch argument depends on the step, so the pair can be captured in a CUDA graph.  Two key buffers suffice:
// a rank can only start writing step e + 2 after its own draft kernel of step e + 1, which waited for every peer's
// look-up of step e + 1, which those peers issued after their draft kernels of step e had consumed the buffer.
#include "samd_common.cuh"
#include "../../include/samd_b200.h"

#include <vector>

#define XCHG_MAX_WORLD 16

struct samd_xchg_s {
    int rank, world, n_queries, device;
    size_t bytes;
    char *local;                                 // this rank's buffer
    char *peer[XCHG_MAX_WORLD];                  // every rank's buffer as seen from here (peer[rank] == local)
    bool opened[XCHG_MAX_WORLD];
};

struct XchgView {
    unsigned long long *keys[XCHG_MAX_WORLD];    // [2][Q] on every rank
    int *flags[XCHG_MAX_WORLD];                  // [world] on every rank
    int *epoch, *done;                           // local
    int rank, world, n;
};

static size_t xchg_keys_bytes(int n) { return (size_t)2 * n * sizeof(unsigned long long); }

static XchgView xchg_view(const samd_xchg_s *x) {
    XchgView v;
    for (int g = 0; g < x->world; ++g) {
        v.keys[g] = reinterpret_cast<unsigned long long *>(x->peer[g]);
        v.flags[g] = reinterpret_cast<int *>(x->peer[g] + xchg_keys_bytes(x->n_queries));
    }
    int *tail = reinterpret_cast<int *>(x->local + xchg_keys_bytes(x->n_queries)) + XCHG_MAX_WORLD;
    v.epoch = tail;
    v.done = tail + 1;                           // [3]: look-up kernel, draft kernel, timed-out flag
    v.rank = x->rank, v.world = x->world, v.n = x->n_queries;
    return v;
}

extern "C" int samd_xchg_create(int rank, int world, int n_queries, samd_xchg_t *out) {
    SAMD_REQUIRE(out && world >= 1 && world <= XCHG_MAX_WORLD && rank >= 0 && rank < world && n_queries > 0,
                 "samd_xchg_create: bad arguments");
    samd_xchg_s *x = new samd_xchg_s();
    x->rank = rank, x->world = world, x->n_queries = n_queries;
    SAMD_CUDA(cudaGetDevice(&x->device));
    x->bytes = xchg_keys_bytes(n_queries) + (XCHG_MAX_WORLD + 4) * sizeof(int);
    SAMD_CUDA(cudaMalloc(&x->local, x->bytes));
    SAMD_CUDA(cudaMemset(x->local, 0, x->bytes));
    SAMD_CUDA(cudaDeviceSynchronize());
    for (int g = 0; g < XCHG_MAX_WORLD; ++g) x->peer[g] = nullptr, x->opened[g] = false;
    x->peer[rank] = x->local;
    *out = x;
    return 0;
}

extern "C" int samd_xchg_export(samd_xchg_t x, void *handle64_out) {
    SAMD_REQUIRE(x && handle64_out, "samd_xchg_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SAMD_CUDA(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle64_out, &h, sizeof(h));
    return 0;
}

extern "C" int samd_xchg_connect(samd_xchg_t x, const void *handles) {
    SAMD_REQUIRE(x && handles, "samd_xchg_connect: bad arguments");
    for (int g = 0; g < x->world; ++g) {
        if (g == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)g * sizeof(h), sizeof(h));
        void *p = nullptr;
        SAMD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[g] = static_cast<char *>(p);
        x->opened[g] = true;
    }
    return 0;
}

extern "C" int samd_xchg_status(samd_xchg_t x) {               // 0 ok, 1 a wait for a peer timed out (synchronises)
    if (!x) return 2;
    int flag = 0;
    XchgView v = xchg_view(x);
    if (cudaMemcpy(&flag, v.done + 2, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
    return flag;
}

extern "C" int samd_xchg_destroy(samd_xchg_t x) {
    if (!x) return 0;
    for (int g = 0; g < x->world; ++g)
        if (x->opened[g]) cudaIpcCloseMemHandle(x->peer[g]);
    cudaFree(x->local);
    delete x;
    return 0;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) xchg_lookup_kernel(StaticDev st, int32_t *cursor, const int32_t *tokens, int stride,
                                                         const int32_t *counts, const int32_t *start_tok, long long shard_offset,
                                                         XchgView X) {
    const int r = blockIdx.x, lane = threadIdx.x;
    const int epoch = *reinterpret_cast<volatile int *>(X.epoch) + 1;      // bumped by the draft kernel's last block
    const int par = epoch & 1;
    int idx = cursor[2 * r], len = cursor[2 * r + 1];
    int hops = 0;
    if (tokens) {                                                          // StaticSAM.transfer_tokens first (static_sam.py:102-104)
        const int k = counts ? samd_clamp_count(counts[r], stride) : stride;
        const int32_t *tk = tokens + (size_t)r * stride;
        for (int i = 0; i < k; i += 32) {
            const int mine = (i + lane < k) ? tk[i + lane] : 0;
            const int lim = min(32, k - i);
            for (int j = 0; j < lim; ++j)
                warp_transfer<true>(st.recs, st.slots, st.bmask, idx, len, __shfl_sync(SAMD_FULL, mine, j), lane, hops);
        }
        if (lane == 0) {
            cursor[2 * r] = idx;
            cursor[2 * r + 1] = len;
        }
    }
    warp_transfer<true>(st.recs, st.slots, st.bmask, idx, len, start_tok[r], lane, hops);
    unsigned long long key = 0;
    if (len > 0) {
        const long long e = shard_offset + (long long)__ldg(st.recs + (size_t)idx * SAMD_REC + R_END);
        key = ((unsigned long long)len << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)e);
    }
    // one atomic per rank, remote over NVLink for the peers: system scope, the only scope that is defined across GPUs
    if (key != 0 && lane < X.world) atomicMax_system(X.keys[lane] + (size_t)par * X.n + r, key);
    __syncwarp();
    if (lane == 0) {
        __threadfence_system();                                               // this block's atomics, before it counts
        if (atomicAdd(&X.done[0], 1) == (int)gridDim.x - 1) {
            X.done[0] = 0;
            __threadfence_system();
            for (int g = 0; g < X.world; ++g) *reinterpret_cast<volatile int *>(&X.flags[g][X.rank]) = epoch;
        }
    }
}

__global__ void xchg_draft_kernel(XchgView X, const int32_t *corpus, long long n_tokens, const int32_t *start_tok, int n_predicts,
                                  int32_t *out_match, int32_t *out_draft, int stride) {
    const int epoch = *reinterpret_cast<volatile int *>(X.epoch) + 1;
    const int par = epoch & 1;
    if (threadIdx.x < X.world) {                                              // every shard's keys have landed here
        const volatile int *f = X.flags[X.rank] + threadIdx.x;
        unsigned long long t0 = 0, t = 0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        // flags only ever grow, and a peer may already be one step ahead (its look-up of step e+1 can finish before this
        // kernel starts polling; the two key buffers cover exactly that skew): wait until the flag has REACHED the epoch
        while ((int)(*f - epoch) < 0) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > 5000000000ull) {                                     // a peer is gone: give up after 5 s, flag it
                X.done[2] = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (r < X.n) {
        unsigned long long *kp = X.keys[X.rank] + (size_t)par * X.n + r;
        unsigned long long key = 0;
        if (lane == 0) {
            key = __ldcg(kp);
            *kp = 0;                                                          // re-armed for the step after next
        }
        key = __shfl_sync(SAMD_FULL, key, 0);
        const int len = (int)(key >> 32);
        const long long e = len > 0 ? (long long)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)) : 0;
        if (lane == 0 && out_match) out_match[r] = len;
        for (int j = lane; j < stride; j += 32) {
            int v = 0;
            if (j < n_predicts) {
                if (j == 0) v = start_tok[r];
                else {
                    const long long pos = e + j;
                    v = pos <= n_tokens ? corpus[pos] : 0;
                }
            }
            out_draft[(size_t)r * stride + j] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&X.done[1], 1) == (int)gridDim.x - 1) {                 // last block out publishes the epoch
            X.done[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile int *>(X.epoch) = epoch;
        }
    }
}

extern "C" int samd_static_lookup_exchange(samd_static_t h, int32_t *static_cursor_dev, const int32_t *tokens_dev,
                                           int32_t token_stride, const int32_t *counts_dev, const int32_t *start_tok_dev,
                                           int64_t shard_offset, samd_xchg_t x, void *stream) {
    SAMD_REQUIRE(h && h->dev.recs && static_cursor_dev && start_tok_dev && x, "samd_static_lookup_exchange: bad arguments");
    SAMD_REQUIRE(!tokens_dev || token_stride > 0, "samd_static_lookup_exchange: token_stride must be positive");
    for (int g = 0; g < x->world; ++g) SAMD_REQUIRE(x->peer[g], "samd_static_lookup_exchange: peers are not connected");
    xchg_lookup_kernel<<<x->n_queries, 32, 0, (cudaStream_t)stream>>>(h->dev, static_cursor_dev, tokens_dev, token_stride, counts_dev,
                                                                      start_tok_dev, (long long)shard_offset, xchg_view(x));
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int samd_draft_from_exchange(samd_xchg_t x, const int32_t *corpus_dev, int64_t n_corpus_tokens,
                                        const int32_t *start_tok_dev, int32_t n_predicts, int32_t *out_match_dev,
                                        int32_t *out_draft_dev, int32_t draft_stride, void *stream) {
    SAMD_REQUIRE(x && corpus_dev && start_tok_dev && out_draft_dev && draft_stride >= n_predicts,
                 "samd_draft_from_exchange: bad arguments");
    const int wpb = 8;
    xchg_draft_kernel<<<(x->n_queries + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        xchg_view(x), corpus_dev, (long long)n_corpus_tokens, start_tok_dev, n_predicts, out_match_dev, out_draft_dev, draft_stride);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
