// Fused greedy tree verification + KV-cache compaction, one persistent launch (sm_100a).
//   phase 1  stream logits [B,T,V] once: per (request,row,chunk) work item a CTA computes the
//            chunk argmax (torch.argmax rule: lowest index among equal maxima, NaN is the
//            maximum, +0 == -0) and folds it into a packed 64-bit key per row with atomicMax.
//            The CTA that completes a request's last item walks the P x D path table
//            (samd/utils.py:127-141), writes the outputs, snapshots + bumps cache_len and
//            publishes the request as ready.
//   phase 2  KV work items (request, tensor): rows cache_len+indices[j] -> cache_len+j
//            (samd/cache.py:118-133), all heads, 16-byte columns, loads before stores.
#include "samd_common.cuh"
#include "../../include/samd_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#define VT 256                 // threads per CTA
#define KV_GROUP 8             // accepted rows staged in registers per pass

struct samd_verify_s {
    unsigned long long *node_key;   // [max_batch][max_nodes]
    int *done;                      // [max_batch]
    int *ready;                     // [max_batch] epoch of the last finished walk
    int *kv_start;                  // [max_batch] cache_len before the bump
    int *work;                      // [4] {phase-1 counter, phase-2 counter, exit counter, unused}
    int max_batch, max_nodes, epoch, device, n_sms;
};

struct VerifyParams {
    samd_verify_args a;
    unsigned long long *node_key;
    int *done, *ready, *kv_start, *work;
    int max_nodes, epoch;
    int chunk, chunks_per_row, n_items1, n_items2, vec_ok;
};

template <int kDtype>
__device__ __forceinline__ uint32_t orderable16(uint32_t b) {
    const uint32_t nan_above = kDtype == SAMD_DTYPE_BF16 ? 0x7F80u : 0x7C00u;
    const uint32_t a = b & 0x7FFFu;
    if (a > nan_above) return 0xFFFFu;            // any NaN is the maximum
    if (a == 0) return 0x8000u;                   // +0 == -0
    return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}

template <int kDtype>
__device__ __forceinline__ uint32_t vec_max_bits(const uint4 &x) {
    if (kDtype == SAMD_DTYPE_BF16) {
        __nv_bfloat162 a = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162 *>(&x.x),
                                       *reinterpret_cast<const __nv_bfloat162 *>(&x.y));
        __nv_bfloat162 b = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162 *>(&x.z),
                                       *reinterpret_cast<const __nv_bfloat162 *>(&x.w));
        a = __hmax2_nan(a, b);
        __nv_bfloat16 s = __hmax_nan(__low2bfloat16(a), __high2bfloat16(a));
        return (uint32_t)__bfloat16_as_ushort(s);
    } else {
        __half2 a = __hmax2_nan(*reinterpret_cast<const __half2 *>(&x.x), *reinterpret_cast<const __half2 *>(&x.y));
        __half2 b = __hmax2_nan(*reinterpret_cast<const __half2 *>(&x.z), *reinterpret_cast<const __half2 *>(&x.w));
        a = __hmax2_nan(a, b);
        __half s = __hmax_nan(__low2half(a), __high2half(a));
        return (uint32_t)__half_as_ushort(s);
    }
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <int kDtype>
__device__ __forceinline__ void fold_vec(const uint4 &x, uint32_t elem, uint32_t &best_key, uint32_t &best_idx) {
    const uint32_t k = orderable16<kDtype>(vec_max_bits<kDtype>(x));
    if (k > best_key) {                            // rare after the first few vectors
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t ke = orderable16<kDtype>((w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
            if (ke > best_key) {
                best_key = ke;
                best_idx = elem + e;
            }
        }
    }
}

__device__ __forceinline__ int ri_at(const VerifyParams &P, int b, int p, int j) {
    if (!P.a.retrieve_dev) return j;                          // sequence: identity path
    return P.a.retrieve_dev[(size_t)b * P.a.retrieve_batch_stride + (size_t)p * P.a.depth + j];
}

template <int kDtype>
__global__ void __launch_bounds__(VT, 4) verify_compact_kernel(VerifyParams P) {
    __shared__ int s_item;
    __shared__ unsigned long long s_red[VT / 32];
    __shared__ int s_flag;
    __shared__ unsigned int s_best;
    extern __shared__ int s_am[];                              // [n_nodes] node argmax of one request
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const samd_verify_args &A = P.a;
    const int T = A.n_nodes;
    const int C = P.chunks_per_row;
    const uint16_t *logits = reinterpret_cast<const uint16_t *>(A.logits_dev);

    // ------------------------------ phase 1: argmax + path walk ---------------------------
    while (true) {
        if (tid == 0) s_item = atomicAdd(&P.work[0], 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= P.n_items1) break;
        const int c = item % C;
        const int t = (item / C) % T;
        const int b = item / (C * T);
        const int n_rows = A.n_nodes_dev ? A.n_nodes_dev[b] : T;
        if (t < n_rows) {
            const int e0 = c * P.chunk;
            const int len = min(P.chunk, A.vocab - e0);
            const uint16_t *row = logits + (size_t)b * A.batch_stride + (size_t)t * A.row_stride + e0;
            uint32_t best_key = 0, best_idx = 0;
            if (P.vec_ok) {
                const uint4 *v4 = reinterpret_cast<const uint4 *>(row);
                const int nvec = len >> 3;
                for (int v = tid; v < nvec; v += 4 * VT) {
                    uint4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (v + u * VT < nvec) x[u] = ld_stream(v4 + v + u * VT);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (v + u * VT < nvec) fold_vec<kDtype>(x[u], (uint32_t)(e0 + (v + u * VT) * 8), best_key, best_idx);
                }
                const int tail = nvec << 3;
                if (tid < len - tail) {                        // < 8 trailing elements
                    const uint32_t ke = orderable16<kDtype>(row[tail + tid]);
                    // a tail element can only beat this thread's vector maxima, never tie with an earlier index
                    if (ke > best_key) {
                        best_key = ke;
                        best_idx = (uint32_t)(e0 + tail + tid);
                    }
                }
            } else {
                for (int e = tid; e < len; e += VT) {
                    const uint32_t ke = orderable16<kDtype>(row[e]);
                    if (ke > best_key) {
                        best_key = ke;
                        best_idx = (uint32_t)(e0 + e);
                    }
                }
            }
            unsigned long long pk = ((unsigned long long)best_key << 32) | (unsigned long long)(0xFFFFFFFFu - best_idx);
            if (best_key == 0) pk = 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(SAMD_FULL, pk, o);
                pk = other > pk ? other : pk;
            }
            if (lane == 0) s_red[warp] = pk;
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < VT / 32; ++w) pk = s_red[w] > pk ? s_red[w] : pk;
                atomicMax(&P.node_key[(size_t)b * P.max_nodes + t], pk);
            }
        }
        if (tid == 0) {
            __threadfence();
            const int prev = atomicAdd(&P.done[b], 1);
            s_flag = (prev == T * C - 1);
        }
        __syncthreads();
        if (!s_flag) continue;

        // ---- this CTA finished request b: path walk (samd/utils.py:127-141) -------------------
        __threadfence();
        if (tid == 0) {
            P.done[b] = 0;
            s_best = 0;
        }
        for (int i = tid; i < T; i += VT) {
            unsigned long long *kp = &P.node_key[(size_t)b * P.max_nodes + i];
            const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(kp);
            const int am = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
            s_am[i] = am;
            *kp = 0;                                            // re-arm for the next launch
            if (A.out_node_argmax_dev && i < n_rows) A.out_node_argmax_dev[(size_t)b * T + i] = am;
        }
        __syncthreads();
        const int32_t *tok = A.tree_tokens_dev + (size_t)b * T;
        const int n_paths = A.retrieve_dev ? (A.n_paths_dev ? A.n_paths_dev[b] : A.n_paths) : 1;
        const int depth = A.retrieve_dev ? A.depth : n_rows;
        for (int p = tid; p < n_paths; p += VT) {
            int acc = 0;
            int prev = ri_at(P, b, p, 0);
            for (int j = 0; j + 1 < depth; ++j) {
                const int nxt = ri_at(P, b, p, j + 1);
                const int rowi = prev < 0 ? n_rows - 1 : prev;           // -1 wraps to the last row
                const int cand = nxt < 0 ? 0 : tok[nxt];                 // -1 selects the appended 0
                if (cand != s_am[rowi]) break;
                acc++;
                prev = nxt;
            }
            atomicMax(&s_best, ((unsigned)acc << 16) | (unsigned)(0xFFFF - p));   // max accept, first path
        }
        __syncthreads();
        const int acc = (int)(s_best >> 16);
        const int best = acc == 0 ? 0 : (int)(0xFFFF - (s_best & 0xFFFF));
        const int out_stride = A.retrieve_dev ? A.depth : T;
        for (int j = tid; j < out_stride; j += VT) {
            int tk = -1, ix = -1;
            if (j <= acc) {
                ix = ri_at(P, b, best, j);
                tk = ix < 0 ? 0 : tok[ix];
            }
            if (A.out_tokens_dev) A.out_tokens_dev[(size_t)b * out_stride + j] = tk;
            if (A.out_indices_dev) A.out_indices_dev[(size_t)b * out_stride + j] = ix;
        }
        if (tid == 0) {
            const int last = ri_at(P, b, best, acc);
            if (A.out_best_dev) A.out_best_dev[b] = best;
            if (A.out_accept_len_dev) A.out_accept_len_dev[b] = acc + 1;
            if (A.out_next_token_dev) A.out_next_token_dev[b] = s_am[last < 0 ? n_rows - 1 : last];
            int start = 0;
            if (A.cache_len_dev) {
                start = A.cache_len_dev[b];
                A.cache_len_dev[b] = start + acc + 1;
            }
            P.kv_start[b] = start;
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicExch(&P.ready[b], P.epoch);
        }
    }

    // ------------------------------ phase 2: KV row moves ---------------------------------
    if (P.n_items2 > 0) {
        const int cols = A.row_bytes >> 4;                      // 16-byte columns per (head,row)
        const int per_tensor = A.n_heads * cols;
        while (true) {
            if (tid == 0) s_item = atomicAdd(&P.work[1], 1);
            __syncthreads();
            const int item = s_item;
            __syncthreads();
            if (item >= P.n_items2) break;
            const int b = item / A.n_kv;
            const int kv = item % A.n_kv;
            if (tid == 0) {
                while (atomicAdd(&P.ready[b], 0) != P.epoch) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
            const int acc1 = __ldcg(A.out_accept_len_dev + b);
            const int start = __ldcg(P.kv_start + b);
            const int32_t *idx = A.out_indices_dev + (size_t)b * A.depth;
            char *base = reinterpret_cast<char *>(__ldg(reinterpret_cast<const unsigned long long *>(A.kv_ptrs_dev) + kv)) +
                         (size_t)b * A.kv_batch_stride;
            for (int j0 = 0; j0 < acc1; j0 += KV_GROUP) {
                int src[KV_GROUP];
#pragma unroll
                for (int u = 0; u < KV_GROUP; ++u) src[u] = (j0 + u < acc1) ? __ldcg(idx + j0 + u) : j0 + u;
                for (int w = tid; w < per_tensor; w += VT) {
                    const int hd = w / cols, col = w - hd * cols;
                    char *hb = base + (size_t)hd * A.kv_head_stride + ((size_t)col << 4);
                    uint4 val[KV_GROUP];
#pragma unroll
                    for (int u = 0; u < KV_GROUP; ++u)
                        if (j0 + u < acc1 && src[u] != j0 + u)
                            val[u] = *reinterpret_cast<const uint4 *>(hb + (size_t)(start + src[u]) * A.kv_pos_stride);
#pragma unroll
                    for (int u = 0; u < KV_GROUP; ++u)
                        if (j0 + u < acc1 && src[u] != j0 + u)
                            *reinterpret_cast<uint4 *>(hb + (size_t)(start + j0 + u) * A.kv_pos_stride) = val[u];
                }
            }
        }
    }
    // last CTA out re-arms the work counters for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&P.work[2], 1) == (int)gridDim.x - 1) {
            P.work[0] = 0;
            P.work[1] = 0;
            P.work[2] = 0;
        }
    }
}

extern "C" int samd_verify_create(int max_batch, int max_nodes, samd_verify_t *out) {
    SAMD_REQUIRE(max_batch > 0 && max_nodes > 0 && out, "samd_verify_create: bad arguments");
    samd_verify_s *h = new samd_verify_s();
    h->max_batch = max_batch;
    h->max_nodes = max_nodes;
    h->epoch = 0;
    SAMD_CUDA(cudaGetDevice(&h->device));
    SAMD_CUDA(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, h->device));
    SAMD_CUDA(cudaMalloc(&h->node_key, (size_t)max_batch * max_nodes * sizeof(unsigned long long)));
    SAMD_CUDA(cudaMalloc(&h->done, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMalloc(&h->ready, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMalloc(&h->kv_start, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMalloc(&h->work, 4 * sizeof(int)));
    SAMD_CUDA(cudaMemset(h->node_key, 0, (size_t)max_batch * max_nodes * sizeof(unsigned long long)));
    SAMD_CUDA(cudaMemset(h->done, 0, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMemset(h->ready, 0, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMemset(h->kv_start, 0, (size_t)max_batch * sizeof(int)));
    SAMD_CUDA(cudaMemset(h->work, 0, 4 * sizeof(int)));
    SAMD_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

extern "C" int samd_verify_destroy(samd_verify_t h) {
    if (!h) return 0;
    cudaFree(h->node_key);
    cudaFree(h->done);
    cudaFree(h->ready);
    cudaFree(h->kv_start);
    cudaFree(h->work);
    delete h;
    return 0;
}

static int g_chunk_override = 0;
extern "C" void samd_verify_set_chunk(int elements) { g_chunk_override = elements; }

extern "C" int samd_verify_compact(samd_verify_t h, const samd_verify_args *a, void *stream) {
    SAMD_REQUIRE(h && a, "samd_verify_compact: bad arguments");
    SAMD_REQUIRE(a->logits_dev && a->tree_tokens_dev, "samd_verify_compact: logits and tree tokens are required");
    SAMD_REQUIRE(a->batch > 0 && a->batch <= h->max_batch, "samd_verify_compact: batch exceeds scratch capacity");
    SAMD_REQUIRE(a->n_nodes > 0 && a->n_nodes <= h->max_nodes, "samd_verify_compact: n_nodes exceeds scratch capacity");
    SAMD_REQUIRE(a->vocab > 0, "samd_verify_compact: bad vocab");
    SAMD_REQUIRE(a->dtype == SAMD_DTYPE_BF16 || a->dtype == SAMD_DTYPE_FP16, "samd_verify_compact: bad dtype");
    SAMD_REQUIRE(!a->retrieve_dev || (a->n_paths > 0 && a->n_paths < 65535 && a->depth > 0),
                 "samd_verify_compact: bad retrieve table shape");
    const bool move = a->move_kv && a->kv_ptrs_dev && a->retrieve_dev;
    if (move) {
        SAMD_REQUIRE(a->out_accept_len_dev && a->out_indices_dev, "samd_verify_compact: KV moves need accept_len and indices outputs");
        SAMD_REQUIRE(a->row_bytes > 0 && a->row_bytes % 16 == 0 && a->kv_pos_stride % 16 == 0 && a->kv_head_stride % 16 == 0 &&
                         a->kv_batch_stride % 16 == 0,
                     "samd_verify_compact: KV rows must be 16-byte aligned");
        SAMD_REQUIRE(a->n_kv > 0 && a->n_heads > 0, "samd_verify_compact: bad KV shape");
    }
    VerifyParams P;
    P.a = *a;
    P.node_key = h->node_key;
    P.done = h->done;
    P.ready = h->ready;
    P.kv_start = h->kv_start;
    P.work = h->work;
    P.max_nodes = h->max_nodes;
    P.epoch = ++h->epoch;
    int chunk = g_chunk_override > 0 ? g_chunk_override : 8192;
    chunk = (chunk + 7) & ~7;
    if (chunk > a->vocab) chunk = (a->vocab + 7) & ~7;
    P.chunk = chunk;
    P.chunks_per_row = (a->vocab + chunk - 1) / chunk;
    P.n_items1 = a->batch * a->n_nodes * P.chunks_per_row;
    P.n_items2 = move ? a->batch * a->n_kv : 0;
    P.vec_ok = ((uintptr_t)a->logits_dev % 16 == 0) && (a->batch_stride % 8 == 0) && (a->row_stride % 8 == 0);
    const size_t smem = (size_t)a->n_nodes * sizeof(int);
    auto kern = a->dtype == SAMD_DTYPE_BF16 ? verify_compact_kernel<SAMD_DTYPE_BF16> : verify_compact_kernel<SAMD_DTYPE_FP16>;
    int per_sm = 0;
    SAMD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, VT, smem));
    SAMD_REQUIRE(per_sm > 0, "samd_verify_compact: kernel does not fit on an SM");
    // persistent grid: every CTA must be resident (phase 2 spins on phase-1 results)
    long long want = (long long)P.n_items1 + P.n_items2;
    int grid = (int)std::min<long long>((long long)h->n_sms * per_sm, want);
    if (grid < 1) grid = 1;
    kern<<<grid, VT, smem, (cudaStream_t)stream>>>(P);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
