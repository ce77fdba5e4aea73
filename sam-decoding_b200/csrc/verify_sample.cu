// Stochastic (typical-acceptance) tree verification - the sampling branch of eval_posterior (samd/utils.py:142-184)
// together with the token draw that follows it (gen_candidates, samd/utils.py:85-88: torch.multinomial(sample_p, 1)).
//
// One CTA per request.  The reference walks the draft level by level: among the paths that share the accepted prefix,
// each distinct candidate token x of level i is tested once, in path order, with one uniform draw r against
//     p(x) / (1 - mass already rejected at this level),        p = softmax(logits_processor(logits[fi, i-1]))
// (logits_processor = temperature, then top-p, then top-k: SamdGenerationConfig.prepare_logits_processor, :44-58); the
// first accepted x extends the prefix, a rejected x is removed from the level's distribution.  The token after the
// accepted prefix is drawn from the level's residual distribution if its last level rejected something, otherwise from the
// plain softmax of the last accepted node's row (no processor - the reference's own asymmetry, :176-179).
//
// RNG contract (samd_b200.h): Philox4x32-10; uniform number k of request b is
//     u = (philox(counter = {lo32(off_b + k), hi32(off_b + k), 0, 0}, key = {lo32(seed_b), hi32(seed_b)})[0] >> 8) * 2^-24
// - one draw per tested candidate, then one for the next token; offsets_dev[b] is advanced by the number of draws used.
// The walk is replicated on every thread of the CTA (same data, same draws): the threads only cooperate on the passes
// over the vocabulary (maximum, the top-p / top-k thresholds by bisection over order-preserving keys, the normaliser, and
// the inverse-CDF draw).
#include "samd_common.cuh"
#include "../../include/samd_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#define ST 256                 // threads per CTA
#define MAX_REJ 64             // distinct rejected tokens remembered per level (>= paths sharing a prefix)

struct PhiloxState {
    uint32_t k0, k1;
    unsigned long long off;
    int used;
};

__device__ __forceinline__ float philox_uniform(PhiloxState &s) {
    const unsigned long long ctr = s.off + (unsigned long long)s.used++;
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
    uint32_t k0 = s.k0, k1 = s.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ uint32_t skey(float f) {          // order-preserving key of a finite float (NaN -> max)
    const uint32_t b = __float_as_uint(f), a = b & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return 0xFFFFFFFFu;
    if (a == 0) return 0x80000000u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int kDtype>
__device__ __forceinline__ float logit_at(const void *base, size_t i) {
    if (kDtype == SAMD_DTYPE_FP32) return __ldg(reinterpret_cast<const float *>(base) + i);
    if (kDtype == SAMD_DTYPE_BF16) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16 *>(base) + i));
    return __half2float(__ldg(reinterpret_cast<const __half *>(base) + i));
}

__device__ __forceinline__ float block_sum(float v, float *s_red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(SAMD_FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < ST / 32; ++w) t += s_red[w];          // same order on every thread: identical result everywhere
    return t;
}
__device__ __forceinline__ float block_max(float v, float *s_red) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SAMD_FULL, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = s_red[0];
    for (int w = 1; w < ST / 32; ++w) t = fmaxf(t, s_red[w]);
    return t;
}

// a row's distribution: p(x) = key(l(x)) >= thr ? exp((l(x) - m) * inv_t) / z : 0
struct RowDist {
    float m, inv_t, z;
    uint32_t thr;
};

template <int kDtype>
__device__ RowDist row_dist(const void *logits, size_t off, int V, bool processed, float temperature, float top_p, int top_k,
                            float *s_red) {
    RowDist d;
    d.inv_t = processed ? 1.0f / temperature : 1.0f;
    d.thr = 0;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < V; i += ST) m = fmaxf(m, logit_at<kDtype>(logits, off + i));
    d.m = block_max(m, s_red);
    auto mass = [&](uint32_t lo_key, bool below) {            // sum of exp over key <= lo_key (below) or key >= lo_key
        float acc = 0.f;
        for (int i = threadIdx.x; i < V; i += ST) {
            const float l = logit_at<kDtype>(logits, off + i);
            const uint32_t k = skey(l);
            if (below ? k <= lo_key : k >= lo_key) acc += __expf((l - d.m) * d.inv_t);
        }
        return block_sum(acc, s_red);
    };
    if (processed) {
        const uint32_t kmax = skey(d.m);
        uint32_t keep_from = 0;                               // keys >= keep_from survive
        if (top_p >= 1e-8f && top_p < 1.0f) {
            // TopPLogitsWarper: ascending cumulative probability <= 1 - top_p is removed (the maximum always stays):
            // the largest key K with mass(key <= K) <= (1 - top_p) * total, found by bisection over the key space
            const float total = mass(0u, false);
            const float budget = (1.0f - top_p) * total;
            uint32_t lo = 0, hi = kmax - 1;                    // invariant: mass(<= lo) <= budget (lo = 0: below every key)
            while (lo < hi) {
                const uint32_t mid = lo + (uint32_t)(((unsigned long long)hi - lo + 1) >> 1);
                if (mass(mid, true) <= budget) lo = mid;
                else hi = mid - 1;
            }
            keep_from = lo + 1;
        }
        if (top_k > 0) {
            // TopKLogitsWarper on what is left: everything below the k-th largest surviving value goes (ties stay)
            auto count_ge = [&](uint32_t key) {
                float c = 0.f;
                for (int i = threadIdx.x; i < V; i += ST) c += skey(logit_at<kDtype>(logits, off + i)) >= key ? 1.f : 0.f;
                return block_sum(c, s_red);
            };
            if (count_ge(keep_from) > (float)top_k) {
                uint32_t lo = keep_from, hi = kmax;            // largest key with count(>= key) >= k
                while (lo < hi) {
                    const uint32_t mid = lo + (uint32_t)(((unsigned long long)hi - lo + 1) >> 1);
                    if (count_ge(mid) >= (float)top_k) lo = mid;
                    else hi = mid - 1;
                }
                keep_from = lo;
            }
        }
        d.thr = keep_from;
    }
    d.z = mass(d.thr, false);
    return d;
}

template <int kDtype>
__device__ __forceinline__ float prob_of(const RowDist &d, const void *logits, size_t off, int x) {
    const float l = logit_at<kDtype>(logits, off + x);
    return skey(l) >= d.thr ? __expf((l - d.m) * d.inv_t) / d.z : 0.f;
}

struct SampleParams {
    samd_sample_args a;
};

template <int kDtype>
__global__ void __launch_bounds__(ST) verify_sample_kernel(SampleParams P) {
    __shared__ float s_red[ST / 32];
    __shared__ float s_part[ST];
    __shared__ int s_rej[MAX_REJ];
    const samd_sample_args &A = P.a;
    const int b = blockIdx.x;
    const int T = A.n_nodes, V = A.vocab;
    const int n_paths = A.retrieve_dev ? (A.n_paths_dev ? A.n_paths_dev[b] : A.n_paths) : 1;
    const int depth = A.retrieve_dev ? A.depth : T;
    const int32_t *tok = A.tree_tokens_dev + (size_t)b * T;
    auto ri = [&](int p, int j) { return A.retrieve_dev ? A.retrieve_dev[(size_t)b * A.retrieve_batch_stride + (size_t)p * A.depth + j] : j; };
    auto cand = [&](int p, int j) {                             // candidate_tokens = tokens_ext[retrieve]: -1 picks the appended 0
        const int r = ri(p, j);
        return r < 0 ? 0 : tok[r];
    };
    auto row_off = [&](int p, int j) {                          // logits[retrieve]: -1 wraps to the last tree row
        const int r = ri(p, j);
        return (size_t)b * A.batch_stride + (size_t)(r < 0 ? T - 1 : r) * A.row_stride;
    };
    PhiloxState rng;
    {
        const unsigned long long seed = A.seeds_dev[b];
        rng.k0 = (uint32_t)seed;
        rng.k1 = (uint32_t)(seed >> 32);
        rng.off = A.offsets_dev[b];
        rng.used = 0;
    }
    int accept_len = 1, best = 0, n_rej = 0;
    bool adjust = false;
    RowDist dist;                                               // the last level's processed distribution
    size_t dist_off = 0;
    for (int i = 1; i < depth; ++i) {
        if (i != accept_len) break;
        adjust = false;
        n_rej = 0;
        // the paths that share the accepted prefix (the prefix is path `best`'s); the first of them supplies the row
        int fi = -1;
        for (int p = 0; p < n_paths && fi < 0; ++p) {
            bool eq = true;
            for (int j = 0; j < accept_len && eq; ++j) eq = cand(p, j) == cand(best, j);
            if (eq) fi = p;
        }
        dist_off = row_off(fi, i - 1);
        dist = row_dist<kDtype>(A.logits_dev, dist_off, V, true, A.temperature, A.top_p, A.top_k, s_red);
        float rej_mass = 0.f;
        const int prefix_of = best;
        for (int p = 0; p < n_paths; ++p) {
            bool eq = true;
            for (int j = 0; j < accept_len && eq; ++j) eq = cand(p, j) == cand(prefix_of, j);
            if (!eq) continue;
            const int x = cand(p, i);
            bool seen = false;
            for (int q = 0; q < n_rej; ++q) seen |= s_rej[q] == x;
            if (seen) continue;
            const float r = philox_uniform(rng);
            const float px = prob_of<kDtype>(dist, A.logits_dev, dist_off, x) / (1.0f - rej_mass);
            if (r <= px) {
                accept_len += 1;
                best = p;
                break;
            }
            rej_mass += prob_of<kDtype>(dist, A.logits_dev, dist_off, x);
            __syncthreads();
            if (threadIdx.x == 0 && n_rej < MAX_REJ) s_rej[n_rej] = x;
            __syncthreads();
            n_rej = min(n_rej + 1, MAX_REJ);
            adjust = true;
        }
    }
    // the distribution of the next token (samd/utils.py:173-179)
    const bool residual = adjust && accept_len != depth;
    if (!residual) {
        dist_off = row_off(best, accept_len - 1);
        dist = row_dist<kDtype>(A.logits_dev, dist_off, V, false, 1.0f, 0.f, 0, s_red);
        n_rej = 0;
    }
    auto p_next = [&](int x) {
        for (int q = 0; q < n_rej; ++q)
            if (s_rej[q] == x) return 0.f;
        return prob_of<kDtype>(dist, A.logits_dev, dist_off, x);
    };
    // inverse-CDF draw in index order: contiguous blocks per thread, a scan over the blocks, then inside the block
    const int chunk = (V + ST - 1) / ST;
    const int x0 = threadIdx.x * chunk, x1 = min(V, x0 + chunk);
    float local = 0.f;
    for (int x = x0; x < x1; ++x) local += p_next(x);
    s_part[threadIdx.x] = local;
    __syncthreads();
    float total = 0.f;
    for (int t = 0; t < ST; ++t) total += s_part[t];
    const float u = philox_uniform(rng) * total;
    float before = 0.f;
    int owner = ST - 1;
    for (int t = 0; t < ST; ++t) {
        if (u < before + s_part[t]) {
            owner = t;
            break;
        }
        before += s_part[t];
    }
    if (A.out_sample_p_dev)
        for (int x = x0; x < x1; ++x) A.out_sample_p_dev[(size_t)b * V + x] = p_next(x) / total;
    if (threadIdx.x == owner) {
        float c = before;
        int pick = -1, last_pos = -1;
        for (int x = x0; x < x1; ++x) {
            const float px = p_next(x);
            if (px > 0.f) last_pos = x;
            c += px;
            if (u < c && px > 0.f) {
                pick = x;
                break;
            }
        }
        if (pick < 0) pick = last_pos >= 0 ? last_pos : max(0, x1 - 1);     // rounding at the very end of the block
        if (A.out_next_token_dev) A.out_next_token_dev[b] = pick;
    }
    if (threadIdx.x == 0) {
        if (A.out_best_dev) A.out_best_dev[b] = best;
        if (A.out_accept_len_dev) A.out_accept_len_dev[b] = accept_len;
        A.offsets_dev[b] = rng.off + (unsigned long long)rng.used;
    }
    const int out_stride = depth;
    for (int j = threadIdx.x; j < out_stride; j += ST) {
        if (A.out_tokens_dev) A.out_tokens_dev[(size_t)b * out_stride + j] = j < accept_len ? cand(best, j) : -1;
        if (A.out_indices_dev) A.out_indices_dev[(size_t)b * out_stride + j] = j < accept_len ? ri(best, j) : -1;
    }
}

extern "C" int samd_verify_sample(const samd_sample_args *a, void *stream) {
    SAMD_REQUIRE(a && a->logits_dev && a->tree_tokens_dev && a->seeds_dev && a->offsets_dev, "samd_verify_sample: bad arguments");
    SAMD_REQUIRE(a->batch > 0 && a->n_nodes > 0 && a->vocab > 0, "samd_verify_sample: bad shape");
    SAMD_REQUIRE(a->temperature >= 1e-5f, "samd_verify_sample: temperature must be >= 1e-5 (samd/utils.py:41)");
    SAMD_REQUIRE(!a->retrieve_dev || (a->n_paths > 0 && a->depth > 0), "samd_verify_sample: bad retrieve table shape");
    SAMD_REQUIRE(a->dtype == SAMD_DTYPE_BF16 || a->dtype == SAMD_DTYPE_FP16 || a->dtype == SAMD_DTYPE_FP32, "samd_verify_sample: bad dtype");
    SampleParams P;
    P.a = *a;
    auto kern = a->dtype == SAMD_DTYPE_BF16 ? verify_sample_kernel<SAMD_DTYPE_BF16>
                : a->dtype == SAMD_DTYPE_FP16 ? verify_sample_kernel<SAMD_DTYPE_FP16> : verify_sample_kernel<SAMD_DTYPE_FP32>;
    kern<<<a->batch, ST, 0, (cudaStream_t)stream>>>(P);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
