// Static suffix automaton: host-side online builder into the flat device layout, upload,
// persistence, L2 window, and the sam_only best-first tree drafter.
#include "samd_common.cuh"
#include "../../include/samd_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---------------------------------------------------------------------------------------
// host builder (StaticSAM.build, samd/sam/static_sam.py:32-79; counts/top-k of
// samd_sam_only/sam/static_sam.py:94-96,137-146) - same record / overflow layout as the device
// ---------------------------------------------------------------------------------------
namespace {

struct HostSam {
    int32_t *recs = nullptr;
    uint4 *slots = nullptr;
    int32_t *text = nullptr;
    uint64_t s_cap = 0, h_cap = 0;
    uint32_t bmask = 0;
    int64_t n_states = 1, n = 0, n_edges = 0, n_clones = 0, n_ovf = 0;
    int last = 0;
    bool growable = false;
    std::vector<uint8_t> is_clone;

    int32_t *rec(int64_t v) const { return recs + (size_t)v * SAMD_REC; }

    // overflow table: found -> slot of the edge; else slot = first free slot of the probe sequence
    bool ovf_find(uint32_t state, uint32_t tok, uint32_t &slot) const {
        uint32_t b = samd_hash(state, tok) & bmask;
        for (;;) {
            const uint4 *bk = slots + (size_t)b * SAMD_BUCKET;
            for (int l = 0; l < SAMD_BUCKET; ++l) {
                if (bk[l].x == state && bk[l].y == tok) {
                    slot = b * SAMD_BUCKET + l;
                    return true;
                }
                if (bk[l].x == SAMD_EMPTY) {
                    slot = b * SAMD_BUCKET + l;
                    return false;
                }
            }
            b = (b + 1) & bmask;
        }
    }
    // pointer to the target word of edge (state, tok), or nullptr
    int32_t *find(int state, int tok) const {
        int32_t *r = rec(state);
        for (int i = 0; i < SAMD_INLINE; ++i) {
            if (r[R_TOK + i] == tok) return r + R_TGT + i;
            if ((uint32_t)r[R_TOK + i] == SAMD_EMPTY) return nullptr;     // inline fills in order, overflow only after
        }
        if ((uint32_t)r[R_OHEAD] == SAMD_NIL) return nullptr;
        uint32_t slot;
        if (!ovf_find((uint32_t)state, (uint32_t)tok, slot)) return nullptr;
        return reinterpret_cast<int32_t *>(&slots[slot].z);
    }
    std::vector<int32_t> ovf_states;                      // states that own an overflow list (for grow())
    // The overflow table doubles when half full (host builder only: the device arenas are sized up front).  Sizing it
    // for the worst case - 4 slots per token - costs 8.6 GB of touched memory for a 125 M-token shard of which a few
    // hundred MB are ever used: only hubs and the root overflow.  Per-state list order is preserved.
    void grow() {
        const uint64_t ncap = h_cap * 2;
        uint4 *nsl = (uint4 *)malloc(ncap * sizeof(uint4));
        if (!nsl) return;                                  // keep filling the old table; the hard check below still holds
        memset(nsl, 0xFF, ncap * sizeof(uint4));
        const uint32_t nmask = (uint32_t)(ncap / SAMD_BUCKET - 1);
        std::vector<uint32_t> map((size_t)h_cap, SAMD_NIL);
        for (uint64_t i = 0; i < h_cap; ++i) {
            if (slots[i].x == SAMD_EMPTY) continue;
            uint32_t bk = samd_hash(slots[i].x, slots[i].y) & nmask, pos = 0;
            for (bool placed = false; !placed; bk = (bk + 1) & nmask)
                for (int l = 0; l < SAMD_BUCKET && !placed; ++l)
                    if (nsl[(size_t)bk * SAMD_BUCKET + l].x == SAMD_EMPTY) {
                        pos = bk * SAMD_BUCKET + l;
                        placed = true;
                    }
            nsl[pos] = slots[i];
            map[(size_t)i] = pos;
        }
        for (uint64_t i = 0; i < h_cap; ++i)
            if (map[(size_t)i] != SAMD_NIL && slots[i].w != SAMD_NIL) nsl[map[(size_t)i]].w = map[slots[i].w];
        for (int32_t v : ovf_states) {
            int32_t *r = rec(v);
            r[R_OHEAD] = (int32_t)map[(uint32_t)r[R_OHEAD]];
            r[R_OTAIL] = (int32_t)map[(uint32_t)r[R_OTAIL]];
        }
        free(slots);
        slots = nsl;
        h_cap = ncap;
        bmask = nmask;
    }
    void add_edge(int state, int tok, int target) {
        int32_t *r = rec(state);
        n_edges++;
        for (int i = 0; i < SAMD_INLINE; ++i)
            if ((uint32_t)r[R_TOK + i] == SAMD_EMPTY) {
                r[R_TOK + i] = tok;
                r[R_TGT + i] = target;
                return;
            }
        if (growable && 2 * ((uint64_t)n_ovf + 1) > h_cap) grow();
        if ((uint32_t)r[R_OHEAD] == SAMD_NIL) ovf_states.push_back(state);
        uint32_t slot;
        ovf_find((uint32_t)state, (uint32_t)tok, slot);
        slots[slot] = make_uint4((uint32_t)state, (uint32_t)tok, (uint32_t)target, SAMD_NIL);
        if ((uint32_t)r[R_OTAIL] != SAMD_NIL) slots[(uint32_t)r[R_OTAIL]].w = slot;
        else r[R_OHEAD] = (int32_t)slot;
        r[R_OTAIL] = (int32_t)slot;
        n_ovf++;
    }
    template <class F>
    void for_each_edge(int64_t v, F f) const {
        const int32_t *r = rec(v);
        for (int i = 0; i < SAMD_INLINE && (uint32_t)r[R_TOK + i] != SAMD_EMPTY; ++i) f(r[R_TOK + i], r[R_TGT + i]);
        for (uint32_t e = (uint32_t)r[R_OHEAD]; e != SAMD_NIL; e = slots[e].w) f((int32_t)slots[e].y, (int32_t)slots[e].z);
    }
    void append(int tok) {
        n += 1;
        const int cur = (int)n_states++;
        samd_init_rec(rec(cur), -1, (int)n, (int)n);
        is_clone.push_back(0);
        text[n] = tok;
        int p = last;
        int32_t *hit = nullptr;
        while (p != -1 && !(hit = find(p, tok))) {
            add_edge(p, tok, cur);
            p = rec(p)[R_LINK];
        }
        if (p == -1) {
            rec(cur)[R_LINK] = 0;
        } else {
            const int q = *hit;
            if (rec(p)[R_LEN] + 1 == rec(q)[R_LEN]) {
                rec(cur)[R_LINK] = q;
            } else {
                const int clone = (int)n_states++;
                n_clones++;
                is_clone.push_back(1);
                // the clone keeps q's link, min_endpos and edges (in q's insertion order)
                samd_init_rec(rec(clone), rec(q)[R_LINK], rec(p)[R_LEN] + 1, rec(q)[R_END]);
                for_each_edge(q, [&](int32_t t, int32_t g) { add_edge(clone, t, g); });
                while (p != -1 && (hit = find(p, tok)) && *hit == q) {
                    *hit = clone;
                    p = rec(p)[R_LINK];
                }
                rec(q)[R_LINK] = clone;
                rec(cur)[R_LINK] = clone;
            }
        }
        last = cur;
    }
};

void free_handle(samd_static_s *h) {
    if (!h) return;
    cudaFree((void *)h->dev.recs);
    cudaFree((void *)h->dev.slots);
    cudaFree((void *)h->dev.text);
    cudaFree((void *)h->dev.occ);
    cudaFree((void *)h->dev.topk);
    free(h->h_recs);
    free(h->h_slots);
    free(h->h_text);
    free(h->h_occ);
    free(h->h_topk);
    delete h;
}

int upload(samd_static_s *h) {
    StaticDev &d = h->dev;
    SAMD_CUDA(cudaGetDevice(&h->device));
    void *p = nullptr;
    SAMD_CUDA(cudaMalloc(&p, (size_t)d.n_states * SAMD_REC * sizeof(int32_t)));
    d.recs = (const int32_t *)p;
    SAMD_CUDA(cudaMemcpy(p, h->h_recs, (size_t)d.n_states * SAMD_REC * sizeof(int32_t), cudaMemcpyHostToDevice));
    SAMD_CUDA(cudaMalloc(&p, (size_t)d.n_slots * sizeof(uint4)));
    d.slots = (const uint4 *)p;
    SAMD_CUDA(cudaMemcpy(p, h->h_slots, (size_t)d.n_slots * sizeof(uint4), cudaMemcpyHostToDevice));
    SAMD_CUDA(cudaMalloc(&p, (size_t)(d.n_tokens + 1) * sizeof(int32_t)));
    d.text = (const int32_t *)p;
    SAMD_CUDA(cudaMemcpy(p, h->h_text, (size_t)(d.n_tokens + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (h->with_counts) {
        SAMD_CUDA(cudaMalloc(&p, (size_t)d.n_states * sizeof(int32_t)));
        d.occ = (const int32_t *)p;
        SAMD_CUDA(cudaMemcpy(p, h->h_occ, (size_t)d.n_states * sizeof(int32_t), cudaMemcpyHostToDevice));
        SAMD_CUDA(cudaMalloc(&p, (size_t)d.n_states * 8 * sizeof(int2)));
        d.topk = (const int2 *)p;
        SAMD_CUDA(cudaMemcpy(p, h->h_topk, (size_t)d.n_states * 8 * sizeof(int2), cudaMemcpyHostToDevice));
    }
    return 0;
}

// Move a finished HostSam into a handle: records trimmed, overflow table re-hashed to its final
// (small) power-of-two size, optional top-k table from the counts.
samd_static_s *finish(HostSam &b, const int32_t *counts, bool with_counts) {
    samd_static_s *h = new samd_static_s();
    memset(h, 0, sizeof(*h));
    const int64_t ns = b.n_states;
    uint64_t cap = samd_next_pow2(2 * (uint64_t)b.n_ovf + 64);
    uint4 *nsl = (uint4 *)malloc(cap * sizeof(uint4));
    memset(nsl, 0xFF, cap * sizeof(uint4));
    const uint32_t nmask = (uint32_t)(cap / SAMD_BUCKET - 1);
    for (int64_t v = 0; v < ns; ++v) {
        int32_t *r = b.rec(v);
        uint32_t head = SAMD_NIL, tail = SAMD_NIL;
        for (uint32_t e = (uint32_t)r[R_OHEAD]; e != SAMD_NIL; e = b.slots[e].w) {
            const uint4 ed = b.slots[e];
            uint32_t bk = samd_hash((uint32_t)v, ed.y) & nmask, slot = 0;
            for (bool placed = false; !placed; bk = (bk + 1) & nmask)
                for (int l = 0; l < SAMD_BUCKET && !placed; ++l)
                    if (nsl[(size_t)bk * SAMD_BUCKET + l].x == SAMD_EMPTY) {
                        slot = bk * SAMD_BUCKET + l;
                        placed = true;
                    }
            nsl[slot] = make_uint4((uint32_t)v, ed.y, ed.z, SAMD_NIL);
            if (tail != SAMD_NIL) nsl[tail].w = slot;
            else head = slot;
            tail = slot;
        }
        r[R_OHEAD] = (int32_t)head;
        r[R_OTAIL] = (int32_t)tail;
    }
    free(b.slots);
    b.slots = nsl;
    b.h_cap = cap;
    b.bmask = nmask;
    h->h_recs = (int32_t *)realloc(b.recs, (size_t)ns * SAMD_REC * sizeof(int32_t));
    b.recs = h->h_recs;
    h->h_slots = nsl;
    h->h_text = b.text;
    h->n_edges = b.n_edges;
    h->n_clones = b.n_clones;
    h->n_ovf = b.n_ovf;
    h->with_counts = with_counts;
    h->dev.n_states = ns;
    h->dev.n_slots = (int64_t)cap;
    h->dev.n_tokens = b.n;
    h->dev.bmask = nmask;
    if (with_counts) {
        h->h_occ = (int32_t *)calloc((size_t)ns, sizeof(int32_t));
        if (counts) {
            memcpy(h->h_occ, counts, (size_t)ns * sizeof(int32_t));
        } else {
            // |endpos| by accumulation up the link tree in decreasing-length order (counting sort); equals the
            // reference's running count (clone inherits q's count, then every state on the new suffix path +1)
            std::vector<int32_t> order((size_t)ns), bucket((size_t)b.n + 2, 0);
            for (int64_t v = 0; v < ns; ++v) bucket[(size_t)b.rec(v)[R_LEN] + 1]++;
            for (size_t i = 1; i < bucket.size(); ++i) bucket[i] += bucket[i - 1];
            for (int64_t v = 0; v < ns; ++v) order[(size_t)bucket[(size_t)b.rec(v)[R_LEN]]++] = (int32_t)v;
            for (int64_t v = 1; v < ns; ++v) h->h_occ[v] = b.is_clone[(size_t)v] ? 0 : 1;
            for (int64_t i = ns - 1; i > 0; --i) {
                const int v = order[(size_t)i];
                const int l = b.rec(v)[R_LINK];
                if (l > 0) h->h_occ[l] += h->h_occ[v];
            }
        }
        h->h_topk = (int2 *)malloc((size_t)ns * 8 * sizeof(int2));
        std::vector<int2> edges;
        for (int64_t v = 0; v < ns; ++v) {
            edges.clear();
            b.for_each_edge(v, [&](int32_t t, int32_t g) { edges.push_back(make_int2(t, g)); });   // dict insertion order
            std::stable_sort(edges.begin(), edges.end(),
                             [&](const int2 &x, const int2 &y) { return h->h_occ[x.y] > h->h_occ[y.y]; });
            for (int j = 0; j < 8; ++j)
                h->h_topk[(size_t)v * 8 + j] = j < (int)edges.size() ? edges[(size_t)j] : make_int2(-1, -1);
        }
    }
    return h;
}

int alloc_host(HostSam &b, uint64_t n_states_cap, uint64_t n_tokens, uint64_t ovf_cap) {
    b.s_cap = n_states_cap;
    b.h_cap = ovf_cap;
    b.bmask = (uint32_t)(b.h_cap / SAMD_BUCKET - 1);
    b.recs = (int32_t *)malloc(b.s_cap * SAMD_REC * sizeof(int32_t));
    b.slots = (uint4 *)malloc(b.h_cap * sizeof(uint4));
    b.text = (int32_t *)calloc((size_t)n_tokens + 1, sizeof(int32_t));
    SAMD_REQUIRE(b.recs && b.slots && b.text, "static SAM: host allocation failed");
    memset(b.slots, 0xFF, b.h_cap * sizeof(uint4));
    samd_init_rec(b.rec(0), -1, 0, 0);
    b.text[0] = -1;
    return 0;
}

}  // namespace

extern "C" int samd_static_build_host(const int32_t *docs, const int64_t *offs, int64_t n_docs, int32_t eos, int with_counts,
                                      samd_static_t *out) {
    SAMD_REQUIRE(docs && offs && n_docs > 0 && out, "samd_static_build: bad arguments");
    int64_t total = 0;
    for (int64_t d = 0; d < n_docs; ++d) {
        const int64_t len = offs[d + 1] - offs[d];
        SAMD_REQUIRE(len > 0, "samd_static_build: empty document");
        total += len + (docs[offs[d + 1] - 1] != eos ? 1 : 0);
    }
    SAMD_REQUIRE(total < (int64_t)1 << 30, "samd_static_build: corpus too large for one shard (>= 2^30 tokens)");
    HostSam b;
    // overflow edges are rare (only states with > 5 out-edges): start at n/2 slots... but never re-hash
    // mid-build, so size for the worst case the corpus allows: edges <= 3n, all of them could overflow
    // only in adversarial inputs; n slots (load <= ~0.5 in practice) with a hard check below.
    // the overflow table starts small and doubles when half full (HostSam::grow): only hubs and the root overflow
    int rc = alloc_host(b, 2 * (uint64_t)total + 2, (uint64_t)total, std::min<uint64_t>(samd_table_slots((uint64_t)total), 1u << 16));
    if (rc) return rc;
    b.growable = true;
    b.is_clone.reserve(b.s_cap);
    b.is_clone.push_back(0);
    for (int64_t d = 0; d < n_docs; ++d) {
        for (int64_t i = offs[d]; i < offs[d + 1]; ++i) {
            SAMD_REQUIRE(docs[i] >= 0, "samd_static_build: token ids must be non-negative");
            b.append(docs[i]);
        }
        if (docs[offs[d + 1] - 1] != eos) b.append(eos);
    }
    *out = finish(b, nullptr, with_counts != 0);
    return 0;
}

// Free the host mirrors of an uploaded automaton (they only serve export / save): a 125 M-token shard keeps 15 GB of them.
extern "C" int samd_static_drop_host(samd_static_t h) {
    SAMD_REQUIRE(h, "samd_static_drop_host: null handle");
    SAMD_REQUIRE(h->dev.recs, "samd_static_drop_host: upload the automaton first");
    free(h->h_recs);
    free(h->h_slots);
    free(h->h_text);
    free(h->h_occ);
    free(h->h_topk);
    h->h_recs = nullptr;
    h->h_slots = nullptr;
    h->h_text = nullptr;
    h->h_occ = nullptr;
    h->h_topk = nullptr;
    return 0;
}

extern "C" int samd_static_upload(samd_static_t h) {
    SAMD_REQUIRE(h && h->h_recs, "samd_static_upload: no host automaton");
    SAMD_REQUIRE(!h->dev.recs, "samd_static_upload: already on the device");
    return upload(h);
}

extern "C" int samd_static_build(const int32_t *docs, const int64_t *offs, int64_t n_docs, int32_t eos, int with_counts,
                                 samd_static_t *out) {
    samd_static_s *h = nullptr;
    int rc = samd_static_build_host(docs, offs, n_docs, eos, with_counts, &h);
    if (rc) return rc;
    if (upload(h)) {
        free_handle(h);
        return 1;
    }
    *out = h;
    return 0;
}

// Converter for automata built elsewhere (reference pickles, samd/sam/utils.py:24-37): explicit
// state arrays plus edges as (state, token, target) triples, per state in dict insertion order.
extern "C" int samd_static_from_arrays(int64_t n_states, const int32_t *link, const int32_t *length, const int32_t *endpos,
                                       const int32_t *count, int64_t n_edges, const int32_t *edges, int64_t n_tokens,
                                       const int32_t *text, samd_static_t *out) {
    SAMD_REQUIRE(n_states > 0 && link && length && n_edges >= 0 && (n_edges == 0 || edges) && out,
                 "samd_static_from_arrays: bad arguments");
    SAMD_REQUIRE(endpos || count, "samd_static_from_arrays: need min_endpos (samd) or cnt_endpos (samd_sam_only)");
    HostSam b;
    int rc = alloc_host(b, (uint64_t)n_states, (uint64_t)n_tokens, samd_next_pow2(2 * (uint64_t)n_edges + 64));
    if (rc) return rc;
    for (int64_t v = 0; v < n_states; ++v) samd_init_rec(b.rec(v), link[v], length[v], endpos ? endpos[v] : 0);
    if (text)
        for (int64_t i = 1; i <= n_tokens; ++i) b.text[i] = text[i];
    for (int64_t e = 0; e < n_edges; ++e) b.add_edge(edges[3 * e], edges[3 * e + 1], edges[3 * e + 2]);
    b.n_states = n_states;
    b.n = n_tokens;
    *out = finish(b, count, count != nullptr);
    return 0;
}

// edges as (state, token, target) triples, per state oldest-first; text[0..n_tokens]
extern "C" int samd_static_export_edges(samd_static_t h, int32_t *edges_host, int32_t *text_host) {
    SAMD_REQUIRE(h && h->h_recs, "samd_static_export_edges: no host mirror");
    if (edges_host) {
        HostSam b;
        b.recs = h->h_recs;
        b.slots = h->h_slots;
        int64_t k = 0;
        for (int64_t v = 0; v < h->dev.n_states; ++v)
            b.for_each_edge(v, [&](int32_t t, int32_t g) {
                edges_host[3 * k] = (int32_t)v;
                edges_host[3 * k + 1] = t;
                edges_host[3 * k + 2] = g;
                k++;
            });
    }
    if (text_host) memcpy(text_host, h->h_text, (size_t)(h->dev.n_tokens + 1) * sizeof(int32_t));
    return 0;
}

extern "C" int samd_static_destroy(samd_static_t h) {
    free_handle(h);
    return 0;
}

extern "C" int samd_static_info(samd_static_t h, int64_t *info) {
    SAMD_REQUIRE(h && info, "samd_static_info: bad arguments");
    info[0] = h->dev.n_states;
    info[1] = h->n_edges;
    info[2] = h->dev.n_tokens;
    info[3] = h->dev.n_slots;
    info[4] = h->dev.n_states * 64 + h->dev.n_slots * 16 + (h->dev.n_tokens + 1) * 4 +
              (h->with_counts ? h->dev.n_states * (4 + 64) : 0);
    info[5] = h->with_counts;
    info[6] = h->n_clones;
    info[7] = h->n_ovf;
    return 0;
}

extern "C" int samd_static_export(samd_static_t h, int32_t *link, int32_t *length, int32_t *endpos, int32_t *count,
                                  int32_t *topk) {
    SAMD_REQUIRE(h && h->h_recs, "samd_static_export: no host mirror");
    for (int64_t v = 0; v < h->dev.n_states; ++v) {
        const int32_t *r = h->h_recs + (size_t)v * SAMD_REC;
        if (link) link[v] = r[R_LINK];
        if (length) length[v] = r[R_LEN];
        if (endpos) endpos[v] = r[R_END];
    }
    if (count) {
        SAMD_REQUIRE(h->h_occ, "samd_static_export: automaton built without counts");
        memcpy(count, h->h_occ, (size_t)h->dev.n_states * sizeof(int32_t));
    }
    if (topk) {
        SAMD_REQUIRE(h->h_topk, "samd_static_export: automaton built without counts");
        memcpy(topk, h->h_topk, (size_t)h->dev.n_states * 8 * sizeof(int2));
    }
    return 0;
}

// flat file: header (8 x int64) then records, overflow slots, text, [occ, topk]
static const int64_t kMagic = 0x30304232444d4153ll;   // "SAMD2B00"
static const int64_t kFormat = 2;                      // 64-byte records with inline edges

extern "C" int samd_static_save(samd_static_t h, const char *path) {
    SAMD_REQUIRE(h && path && h->h_recs, "samd_static_save: bad arguments");
    FILE *f = fopen(path, "wb");
    SAMD_REQUIRE(f, "samd_static_save: cannot open file");
    int64_t hdr[8] = {kMagic, kFormat, h->dev.n_states, h->dev.n_slots, h->dev.n_tokens, h->n_edges,
                      h->with_counts | (h->n_ovf << 8), h->n_clones};
    bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1;
    ok = ok && fwrite(h->h_recs, SAMD_REC * sizeof(int32_t), (size_t)h->dev.n_states, f) == (size_t)h->dev.n_states;
    ok = ok && fwrite(h->h_slots, sizeof(uint4), (size_t)h->dev.n_slots, f) == (size_t)h->dev.n_slots;
    ok = ok && fwrite(h->h_text, sizeof(int32_t), (size_t)h->dev.n_tokens + 1, f) == (size_t)h->dev.n_tokens + 1;
    if (h->with_counts) {
        ok = ok && fwrite(h->h_occ, sizeof(int32_t), (size_t)h->dev.n_states, f) == (size_t)h->dev.n_states;
        ok = ok && fwrite(h->h_topk, sizeof(int2), (size_t)h->dev.n_states * 8, f) == (size_t)h->dev.n_states * 8;
    }
    fclose(f);
    SAMD_REQUIRE(ok, "samd_static_save: short write");
    return 0;
}

static int load_impl(const char *path, samd_static_t *out, bool to_device) {
    SAMD_REQUIRE(path && out, "samd_static_load: bad arguments");
    FILE *f = fopen(path, "rb");
    SAMD_REQUIRE(f, "samd_static_load: cannot open file");
    int64_t hdr[8];
    if (fread(hdr, sizeof(hdr), 1, f) != 1 || hdr[0] != kMagic || hdr[1] != kFormat) {
        fclose(f);
        samd_set_error("samd_static_load: not a samd_b200 automaton file (or format mismatch)");
        return 2;
    }
    samd_static_s *h = new samd_static_s();
    memset(h, 0, sizeof(*h));
    h->dev.n_states = hdr[2];
    h->dev.n_slots = hdr[3];
    h->dev.n_tokens = hdr[4];
    h->n_edges = hdr[5];
    h->with_counts = (int)(hdr[6] & 0xFF);
    h->n_ovf = hdr[6] >> 8;
    h->n_clones = hdr[7];
    h->dev.bmask = (uint32_t)(h->dev.n_slots / SAMD_BUCKET - 1);
    h->h_recs = (int32_t *)malloc((size_t)h->dev.n_states * SAMD_REC * sizeof(int32_t));
    h->h_slots = (uint4 *)malloc((size_t)h->dev.n_slots * sizeof(uint4));
    h->h_text = (int32_t *)malloc((size_t)(h->dev.n_tokens + 1) * sizeof(int32_t));
    bool ok = h->h_recs && h->h_slots && h->h_text;
    ok = ok && fread(h->h_recs, SAMD_REC * sizeof(int32_t), (size_t)h->dev.n_states, f) == (size_t)h->dev.n_states;
    ok = ok && fread(h->h_slots, sizeof(uint4), (size_t)h->dev.n_slots, f) == (size_t)h->dev.n_slots;
    ok = ok && fread(h->h_text, sizeof(int32_t), (size_t)h->dev.n_tokens + 1, f) == (size_t)h->dev.n_tokens + 1;
    if (ok && h->with_counts) {
        h->h_occ = (int32_t *)malloc((size_t)h->dev.n_states * sizeof(int32_t));
        h->h_topk = (int2 *)malloc((size_t)h->dev.n_states * 8 * sizeof(int2));
        ok = h->h_occ && h->h_topk;
        ok = ok && fread(h->h_occ, sizeof(int32_t), (size_t)h->dev.n_states, f) == (size_t)h->dev.n_states;
        ok = ok && fread(h->h_topk, sizeof(int2), (size_t)h->dev.n_states * 8, f) == (size_t)h->dev.n_states * 8;
    }
    fclose(f);
    if (!ok) {
        free_handle(h);
        samd_set_error("samd_static_load: short read");
        return 2;
    }
    if (to_device && upload(h)) {
        free_handle(h);
        return 1;
    }
    *out = h;
    return 0;
}

extern "C" int samd_static_load(const char *path, samd_static_t *out) { return load_impl(path, out, true); }
extern "C" int samd_static_load_host(const char *path, samd_static_t *out) { return load_impl(path, out, false); }

extern "C" int samd_static_set_l2_window(samd_static_t h, void *stream, int64_t bytes) {
    SAMD_REQUIRE(h, "samd_static_set_l2_window: null handle");
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (bytes <= 0) {
        attr.accessPolicyWindow.num_bytes = 0;
        SAMD_CUDA(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        SAMD_CUDA(cudaCtxResetPersistingL2Cache());
        return 0;
    }
    SAMD_REQUIRE(h->dev.recs, "samd_static_set_l2_window: automaton is not on the device");
    cudaDeviceProp prop;
    SAMD_CUDA(cudaGetDeviceProperties(&prop, h->device));
    size_t carve = std::min((size_t)bytes, (size_t)prop.persistingL2CacheMaxSize);
    SAMD_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    // State records are contiguous in creation (= document) order.  The window is the PREFIX that fits the carve-out
    // (root, and the states of the first documents - where the states of short, frequent contexts are created), all
    // of it persisting; everything else streams.
    size_t span = std::min(std::min(carve, (size_t)h->dev.n_states * SAMD_REC * sizeof(int32_t)), (size_t)prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.base_ptr = (void *)h->dev.recs;
    attr.accessPolicyWindow.num_bytes = span;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    SAMD_CUDA(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return 0;
}

// ---------------------------------------------------------------------------------------
// sam_only static tree drafter (samd_sam_only/sam/static_sam.py:148-215)
// one warp per query; lane 0 runs the heap (CPython heapq sift order), the warp builds buffers
// ---------------------------------------------------------------------------------------
#define TREE_MAX 128
#define HEAP_MAX (8 * TREE_MAX + 1)

struct HeapItem {
    double prob;
    int tok, state;
    short parent, depth;
};

__device__ __forceinline__ void heap_sift_down(HeapItem *a, int start, int pos) {   // heapq._siftdown
    HeapItem item = a[pos];
    while (pos > start) {
        const int parent = (pos - 1) >> 1;
        if (item.prob < a[parent].prob) {
            a[pos] = a[parent];
            pos = parent;
            continue;
        }
        break;
    }
    a[pos] = item;
}

__device__ __forceinline__ HeapItem heap_pop(HeapItem *a, int &size) {   // heapq.heappop
    HeapItem last = a[--size];
    if (size == 0) return last;
    HeapItem top = a[0];
    a[0] = last;
    int pos = 0, child = 1;
    while (child < size) {                                                // heapq._siftup
        const int right = child + 1;
        if (right < size && !(a[child].prob < a[right].prob)) child = right;
        a[pos] = a[child];
        pos = child;
        child = 2 * pos + 1;
    }
    a[pos] = last;
    heap_sift_down(a, 0, pos);
    return top;
}

__global__ void __launch_bounds__(32) static_tree_kernel(StaticDev st, int n_req, const int32_t *type, const int32_t *index,
                                                         const int32_t *match, const int32_t *start_tok, int max_predicts,
                                                         double alpha, int K, int len_bias, int32_t *out_tokens,
                                                         int32_t *out_parents, int32_t *out_depth, int32_t *out_n,
                                                         int32_t *out_ret, int max_paths, int max_depth, int32_t *out_shape) {
    // the heap never holds more than K entries per popped node (+ the root): sized by the launch (dynamic shared memory),
    // so that a 40-node draft takes 8 KB instead of the 25 KB of the 128-node worst case - three times the resident queries
    extern __shared__ __align__(16) unsigned char s_heap_raw[];
    HeapItem *heap = reinterpret_cast<HeapItem *>(s_heap_raw);
    __shared__ int s_tok[TREE_MAX], s_par[TREE_MAX], s_dep[TREE_MAX], s_cnt[TREE_MAX], s_leaf[TREE_MAX];
    __shared__ int s_n;
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    if (r >= n_req) return;
    if (type && type[r] != SAMD_DRAFT_STATIC_TREE) {
        if (lane == 0 && out_n) out_n[r] = 0;
        return;
    }
    const int m = match[r] - len_bias;
    const int n = min(min(max_predicts, TREE_MAX), 1 + (int)((double)m * alpha));
    for (int i = lane; i < TREE_MAX; i += 32) s_cnt[i] = 0;
    __syncwarp();
    // Lane 0 owns the heap (the pops and pushes happen in the reference's order, heapq sift for sift); the memory side of
    // an expansion is the warp's: the popped state's top-K row and its K children's occurrence counts are 1 + K
    // dependent-free loads issued side by side (two round trips per node instead of K + 2 - the drafter is bound by the
    // latency of these scattered reads into a multi-gigabyte automaton).
    {
        int size = 0, nt = 0;                                   // lane 0's
        if (lane == 0) heap[size++] = HeapItem{-1.0, start_tok[r], index[r], (short)-1, (short)0};
        while (true) {
            int go = 0, state = 0, me = 0, depth = 0;            // go: 0 = done, 1 = expand the popped node, 2 = popped node dropped
            double prob = 0.0;
            if (lane == 0 && nt != n && size != 0) {
                const HeapItem it = heap_pop(heap, size);
                if (s_cnt[it.depth] + 1 > K) {
                    go = 2;
                } else {
                    s_cnt[it.depth] += 1;
                    me = nt++;
                    s_tok[me] = it.tok;
                    s_par[me] = it.parent;
                    if (nt != n) {
                        go = 1;
                        state = it.state;
                        prob = it.prob;
                        depth = it.depth;
                    }
                }
            }
            go = __shfl_sync(SAMD_FULL, go, 0);
            if (go == 0) break;
            if (go == 2) continue;
            state = __shfl_sync(SAMD_FULL, state, 0);
            const double total = (double)__ldg(st.occ + state);
            int2 e = make_int2(-1, -1);
            int c = 0;
            if (lane < K && lane < 8) {
                e = __ldg(st.topk + (size_t)state * 8 + lane);
                if (e.x >= 0) {
                    c = __ldg(st.occ + e.y);
                    // a child that is popped later starts with its own top-K row: requested now (64 bytes, into L2)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(st.topk + (size_t)e.y * 8));
                }
            }
            // every child's probability is computed by its own lane (an fp64 division is a long dependent sequence; eight of
            // them one after the other were most of an expansion): same operations, same order - division first, then
            // the multiplication - as the reference's `prob * (cnt / total)`
            prob = __shfl_sync(SAMD_FULL, prob, 0);
            const double mine = __dmul_rn(prob, __ddiv_rn((double)c, total));
            for (int j = 0; j < K && j < 8; ++j) {
                const int ex = __shfl_sync(SAMD_FULL, e.x, j), ey = __shfl_sync(SAMD_FULL, e.y, j);
                const double pj = __shfl_sync(SAMD_FULL, mine, j);
                if (ex < 0) break;
                if (lane == 0) {
                    HeapItem ch{pj, ex, ey, (short)me, (short)(depth + 1)};
                    heap[size] = ch;
                    heap_sift_down(heap, 0, size);
                    size++;
                }
            }
        }
        if (lane == 0) s_n = nt;
    }
    __syncwarp();
    const int nt = s_n;
    // gen_buffers (static_sam.py:148-180): depth, leaves ascending, root-to-leaf paths, -1 padding
    for (int i = lane; i < nt; i += 32) s_leaf[i] = 1;
    __syncwarp();
    for (int i = lane; i < nt; i += 32)
        if (i > 0) s_leaf[s_par[i]] = 0;
    for (int i = lane; i < nt; i += 32) {
        int d = 0;
        for (int j = i; s_par[j] >= 0; j = s_par[j]) d++;
        s_dep[i] = d;
    }
    __syncwarp();
    int n_leaves = 0, depth_max = 0;
    for (int base = 0; base < nt; base += 32) {
        const int i = base + lane;
        const bool leaf = i < nt && s_leaf[i];
        const unsigned bal = __ballot_sync(SAMD_FULL, leaf);
        const int rank = n_leaves + __popc(bal & ((1u << lane) - 1));
        int d = leaf ? s_dep[i] + 1 : 0;
        for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(SAMD_FULL, d, o));
        depth_max = max(depth_max, d);
        if (leaf && out_ret && rank < max_paths) {
            int32_t *row = out_ret + ((size_t)r * max_paths + rank) * max_depth;
            for (int c = s_dep[i] + 1; c < max_depth; ++c) row[c] = -1;
            for (int j = i; j >= 0; j = s_par[j])
                if (s_dep[j] < max_depth) row[s_dep[j]] = j;
        }
        n_leaves += __popc(bal);
    }
    for (int i = lane; i < nt; i += 32) {
        if (out_tokens) out_tokens[(size_t)r * max_predicts + i] = s_tok[i];
        if (out_parents) out_parents[(size_t)r * max_predicts + i] = s_par[i];
        if (out_depth) out_depth[(size_t)r * max_predicts + i] = s_dep[i];
    }
    if (lane == 0) {
        if (out_n) out_n[r] = nt;
        if (out_shape) {
            out_shape[2 * r] = n_leaves;
            out_shape[2 * r + 1] = depth_max;
        }
    }
}

extern "C" int samd_static_tree_draft(samd_static_t h, int n_requests, const int32_t *type_dev, const int32_t *index_static_dev,
                                      const int32_t *match_static_dev, const int32_t *start_tok_dev, int32_t max_predicts,
                                      double alpha, int32_t K, int32_t len_bias, int32_t *out_tokens_dev,
                                      int32_t *out_parents_dev, int32_t *out_depth_dev, int32_t *out_n_nodes_dev,
                                      int32_t *out_retrieve_dev, int32_t max_paths, int32_t max_depth,
                                      int32_t *out_retrieve_shape_dev, void *stream) {
    SAMD_REQUIRE(h && h->with_counts, "samd_static_tree_draft: automaton was built without counts");
    SAMD_REQUIRE(n_requests > 0 && index_static_dev && match_static_dev && start_tok_dev, "samd_static_tree_draft: bad arguments");
    SAMD_REQUIRE(max_predicts > 0 && max_predicts <= TREE_MAX, "samd_static_tree_draft: max_predicts must be in [1,128]");
    SAMD_REQUIRE(K > 0 && K <= 8, "samd_static_tree_draft: K must be in [1,8]");
    const size_t heap_bytes = ((size_t)K * (size_t)std::min(max_predicts, TREE_MAX) + 2) * sizeof(HeapItem);
    static_tree_kernel<<<n_requests, 32, heap_bytes, (cudaStream_t)stream>>>(
        h->dev, n_requests, type_dev, index_static_dev, match_static_dev, start_tok_dev, max_predicts, alpha, K, len_bias,
        out_tokens_dev, out_parents_dev, out_depth_dev, out_n_nodes_dev, out_retrieve_dev, max_paths, max_depth,
        out_retrieve_shape_dev);
    samd_count_launch();
    SAMD_CUDA(cudaGetLastError());
    return 0;
}
