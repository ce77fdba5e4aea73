// One-thread formulation of the dynamic suffix automaton's step (DraftModel.update + DraftModel.lookup).
//
// A request's step is ONE serial chain of dependent record reads; what bounds it is the length of that chain in
// instructions and round trips, not bandwidth.  The warp-cooperative probe (sam_step.cu, kernel variant 0) spends a
// ballot, a find-first-set and two or three shuffles between a record's arrival and the address of the next one.
// Here a single thread owns the request: a state record is four 128-bit loads into sixteen registers, a transition
// probe is five compares, the suffix link / length / min_endpos are already in registers when the probe fails, and
// the record of the state the cursor moves to is carried over as the next token's cursor record (no reload).  The
// rest of the warp only joins for the coalesced draft copy.
//
// This header is plain C++ apart from the loads (inline PTX on the device): the same source is compiled for the
// host by tests/host_emul (test infrastructure) so that the append / clone / redirect logic can be checked against
// the oracle on a machine without a GPU.  The product path is the CUDA kernel only.
//
// Reference lines: samd/sam/dyn_sam.py:41-67 (add_state), :69-82 (transfer_state), :84-88 (add_tokens),
// :94-97 (lookup), :99-113 (to_anc, gen_draft); samd_sam_only/sam/dyn_sam.py:116-121.
#pragma once
#include "samd_common.cuh"
#include <string.h>

#if defined(__CUDACC__)
#define SAMD_HD __host__ __device__ __forceinline__
#else
#define SAMD_HD inline
#endif

#define SC_CHAIN_MAX 64

struct Rec {
    int w[SAMD_REC];
};

// ---- memory access ---------------------------------------------------------------------------------------
// kRO: the static automaton (never written while a kernel runs): non-coherent loads.  The dynamic arenas are
// written by the same thread that reads them; the asm statements are volatile with a memory clobber so that the
// compiler keeps them in program order with the plain stores around them.
template <bool kRO>
SAMD_HD Rec rec_load(const int32_t *recs, long long state) {
    Rec r;
    const int32_t *p = recs + (size_t)state * SAMD_REC;
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (kRO)
            asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r.w[4 * i]), "=r"(r.w[4 * i + 1]), "=r"(r.w[4 * i + 2]), "=r"(r.w[4 * i + 3])
                         : "l"(p + 4 * i));
        else
            asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r.w[4 * i]), "=r"(r.w[4 * i + 1]), "=r"(r.w[4 * i + 2]), "=r"(r.w[4 * i + 3])
                         : "l"(p + 4 * i)
                         : "memory");
    }
#else
    memcpy(r.w, p, sizeof(r.w));
#endif
    return r;
}

// the first quarter of a record: {link, length, min_endpos, overflow head}
template <bool kRO>
SAMD_HD int4 rec_head(const int32_t *recs, long long state) {
    int4 v;
    const int32_t *p = recs + (size_t)state * SAMD_REC;
#if defined(__CUDA_ARCH__)
    if (kRO) asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#else
    v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3];
#endif
    return v;
}

template <bool kRO>
SAMD_HD uint4 slot_load(const uint4 *slots, size_t i) {
    uint4 v;
#if defined(__CUDA_ARCH__)
    if (kRO) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slots + i));
    else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slots + i) : "memory");
#else
    v = slots[i];
#endif
    return v;
}

SAMD_HD void rec_store(int32_t *recs, int state, const Rec &r) {
    int4 *p = reinterpret_cast<int4 *>(recs + (size_t)state * SAMD_REC);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_int4(r.w[4 * i], r.w[4 * i + 1], r.w[4 * i + 2], r.w[4 * i + 3]);
}

SAMD_HD void sc_prefetch(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
SAMD_HD void sc_prefetch_rec(const int32_t *recs, long long state) {
    sc_prefetch(recs + (size_t)state * SAMD_REC);
    sc_prefetch(recs + (size_t)state * SAMD_REC + 8);
}

// ---- transition probe ------------------------------------------------------------------------------------
struct Probe {
    int      target;   // transition target when found
    int      k;        // inline index of the hit, else -1
    uint32_t slot;     // overflow slot of the hit, else NIL
    bool     found;
};

// number of inline edges = index of the first free inline slot (they fill in order); SAMD_INLINE when full
SAMD_HD int rec_free_inline(const Rec &X) {
    int f = SAMD_INLINE;
#pragma unroll
    for (int i = SAMD_INLINE - 1; i >= 0; --i)
        if ((uint32_t)X.w[R_TOK + i] == SAMD_EMPTY) f = i;
    return f;
}

// (state, tok) in the overflow table: found -> slot / target; else *free_slot = first free slot of its probe
// sequence.  Nothing is ever deleted, so a probe sequence is a run of used slots followed by a free one.
// `max_buckets` bounds the search of a reader that races with the writer (scouts); 0 = unbounded.
template <bool kRO>
SAMD_HD bool ovf_find(const uint4 *slots, uint32_t bmask, uint32_t state, uint32_t tok, Probe &r, uint32_t *free_slot,
                      int max_buckets = 0) {
    uint32_t b = samd_hash(state, tok) & bmask;
    for (int nb = 0; max_buckets == 0 || nb < max_buckets; ++nb) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const size_t base = (size_t)b * SAMD_BUCKET + half * 4;
            uint4 s[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) s[l] = slot_load<kRO>(slots, base + l);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                if (s[l].x == state && s[l].y == tok) {
                    r.found = true;
                    r.slot = (uint32_t)(base + l);
                    r.target = (int)s[l].z;
                    return true;
                }
                if (s[l].x == SAMD_EMPTY) {
                    if (free_slot) *free_slot = (uint32_t)(base + l);
                    return false;
                }
            }
        }
        b = (b + 1) & bmask;
    }
    return false;
}

template <bool kRO>
SAMD_HD Probe rec_probe(const Rec &X, const uint4 *slots, uint32_t bmask, int state, int tok, int max_buckets = 0) {
    Probe r;
    r.target = 0;
    r.k = -1;
    r.slot = SAMD_NIL;
    r.found = false;
#pragma unroll
    for (int i = 0; i < SAMD_INLINE; ++i)
        if (X.w[R_TOK + i] == tok) {
            r.k = i;
            r.target = X.w[R_TGT + i];
            r.found = true;
        }
    if (!r.found && (uint32_t)X.w[R_OHEAD] != SAMD_NIL) ovf_find<kRO>(slots, bmask, (uint32_t)state, (uint32_t)tok, r, nullptr, max_buckets);
    return r;
}

// ---- the builder -----------------------------------------------------------------------------------------
struct ScDyn {
    int32_t *recs;
    uint4   *slots;
    int32_t *text;
    uint32_t bmask;
    int      max_tokens;
};

// warp-uniform in the old kernel, thread-private here.  ll_*: when the previous append split a state, last_link is
// that clone and `ll_twin` the state it was copied from - their inline edges are identical, so a probe of the twin
// (the cursor's state) answers for the clone as well.
struct ScRegs {
    int n_states, last, last_link, n, cur, cur_len, n_edges, n_clones, hops;
    int ll_twin, ll_len, ll_link;
    int max_chain;
};

struct ScCounters {            // optional per-request trace of the record reads (floor analysis), see tools/
    int32_t *trace;
    int      cap, n;
};

#ifdef SAMD_SCALAR_STATS          // host test build only: which paths of extend_one ran
#define SC_STAT(x) (++stats[x])
#else
#define SC_STAT(x) ((void)0)
#endif
enum { SC_ST_ALIGNED = 0, SC_ST_TWIN = 1, SC_ST_GENERIC = 2, SC_ST_LONGCHAIN = 3, SC_ST_OVF_INSERT = 4, SC_ST_OVF_CLONE = 5,
       SC_ST_CARRIED = 6, SC_ST_N = 8 };

struct ScBuilder {
#ifdef SAMD_SCALAR_STATS
    long long stats[SC_ST_N];
#endif
    ScDyn   d;
    ScRegs  g;
    Rec     X;                 // the cursor's record when x_state == g.cur (carried from token to token)
    int     x_state;
    int    *chain;             // [SC_CHAIN_MAX] states the cursor's walk visited (shared memory on the device)
    unsigned char *cfree;      // [SC_CHAIN_MAX] their first free inline slot at that time
    ScCounters tr;

    SAMD_HD Rec load(int state) {
        if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = state;
        return rec_load<false>(d.recs, state);
    }

    // append (state, tok) -> target to the overflow table and to the state's list (lists run oldest -> newest)
    SAMD_HD void ovf_insert(int state, int tok, int target) {
        Probe dummy;
        uint32_t slot = SAMD_NIL;
        // (state, tok) is known to be absent, so the search ends at the free slot of its probe sequence; fresh
        // reads, so that slots this very append has just filled are seen
        ovf_find<false>(d.slots, d.bmask, (uint32_t)state, (uint32_t)tok, dummy, &slot);
        int32_t *rec = d.recs + (size_t)state * SAMD_REC;
        d.slots[slot] = make_uint4((uint32_t)state, (uint32_t)tok, (uint32_t)target, SAMD_NIL);
        const uint32_t tail = (uint32_t)rec[R_OTAIL];
        if (tail != SAMD_NIL) d.slots[tail].w = slot;
        else rec[R_OHEAD] = (int)slot;
        rec[R_OTAIL] = (int)slot;
    }

    SAMD_HD void insert_edge(int state, int free_inline, int tok, int target) {
        if (free_inline < SAMD_INLINE) {
            int32_t *rec = d.recs + (size_t)state * SAMD_REC;
            rec[R_TOK + free_inline] = tok;
            rec[R_TGT + free_inline] = target;
        } else {
            SC_STAT(SC_ST_OVF_INSERT);
            ovf_insert(state, tok, target);
        }
    }

    // transfer_cur_state + add_state for one token (dyn_sam.py:84-88: match first, then append)
    SAMD_HD void extend_one(int tok) {
        // ---- the cursor's walk (dyn_sam.py:69-78), leaving the states it visits in chain[] ----
        int x = g.cur, len = g.cur_len;
        if (x_state != x) X = load(x);
        else SC_STAT(SC_ST_CARRIED);
        int n_chain = 0;
        bool first = true, on_edge = false;
        Probe pr;
        while (true) {
            pr = rec_probe<false>(X, d.slots, d.bmask, x, tok);
            if (n_chain < SC_CHAIN_MAX) {
                chain[n_chain] = x;
                cfree[n_chain] = (unsigned char)rec_free_inline(X);
            }
            ++n_chain;
            ++g.hops;
            if (!first) len = X.w[R_LEN];                  // length = states[index].length after a link hop
            if (pr.found) {
                on_edge = true;
                break;
            }
            if (x == 0) break;
            x = X.w[R_LINK];
            X = load(x);
            first = false;
        }
        const int new_cur = on_edge ? pr.target : 0;
        const int new_len = on_edge ? len + 1 : 0;
        if (on_edge) sc_prefetch_rec(d.recs, new_cur);      // the record the append reads at the end of its chain

        // ---- add_state (dyn_sam.py:41-67) ----
        g.n += 1;
        const int cur = g.n_states++;
        {
            Rec N;
#pragma unroll
            for (int i = 0; i < SAMD_REC; ++i) N.w[i] = 0;
            N.w[R_LINK] = -1;                               // written at the end
            N.w[R_LEN] = g.n;
            N.w[R_END] = g.n;
            N.w[R_OHEAD] = N.w[R_OTAIL] = -1;
#pragma unroll
            for (int i = 0; i < SAMD_INLINE; ++i) N.w[R_TOK + i] = -1;
            rec_store(d.recs, cur, N);
        }
        d.text[g.n] = tok;
        int p = g.last;
        if (p != 0) {
            // `last` was created by the previous append and has no out-edge yet: its first inline edge needs no read
            d.recs[(size_t)p * SAMD_REC + R_TOK] = tok;
            d.recs[(size_t)p * SAMD_REC + R_TGT] = cur;
            g.n_edges++;
            p = g.last_link;
        }
        // The cursor's walk visits exactly the states the append gives the edge to - the cursor is link(last), or,
        // after a split, the pre-clone state one stop in front of it (sam_step.cu) - so the edge is written at all of
        // them from what the walk saw, without reading anything again.
        int p_len = 0, p_link = -1;
        Probe pp;
        pp.found = false;
        int s0 = -1;
        if (n_chain <= SC_CHAIN_MAX) {
            if (chain[0] == p) s0 = 0;
            else if (n_chain > 1 && chain[1] == p) s0 = 1;
        }
        if (n_chain > SC_CHAIN_MAX) SC_STAT(SC_ST_LONGCHAIN);
        if (n_chain > g.max_chain) g.max_chain = n_chain;
        if (s0 >= 0) {
            SC_STAT(SC_ST_ALIGNED);
            const int end = on_edge ? n_chain - 1 : n_chain;           // [s0, end): states without the edge
            for (int j = s0; j < end; ++j) insert_edge(chain[j], cfree[j], tok, cur);
            if (end > s0) g.n_edges += end - s0;
            if (on_edge) {
                p = x;
                p_len = X.w[R_LEN];
                p_link = X.w[R_LINK];
                pp = pr;
            } else {
                p = -1;
            }
        } else if (on_edge && n_chain == 1 && p >= 0 && g.ll_twin == chain[0] && pr.k >= 0) {
            // last_link is the clone the previous append made of the cursor's state: same inline edges, so the
            // probe of the cursor's record stands for it
            SC_STAT(SC_ST_TWIN);
            p_len = g.ll_len;
            p_link = g.ll_link;
            pp = pr;
        } else {
            SC_STAT(SC_ST_GENERIC);
            while (p != -1) {
                const Rec P = load(p);
                const Probe q = rec_probe<false>(P, d.slots, d.bmask, p, tok);
                if (!q.found) {
                    insert_edge(p, rec_free_inline(P), tok, cur);
                    g.n_edges++;
                    p = P.w[R_LINK];
                    continue;
                }
                p_len = P.w[R_LEN];
                p_link = P.w[R_LINK];
                pp = q;
                break;
            }
        }
        int link_cur = 0;
        g.ll_twin = -1;
        x_state = -1;
        if (p != -1) {
            const int q = pp.target;
            Rec Q = load(q);                                // after every store above (q may be `last` or a chain state)
            if (p_len + 1 == Q.w[R_LEN]) {
                link_cur = q;
            } else {
                // clone-on-split: q's record with length len(p)+1; overflow edges re-inserted oldest first
                const int clone = g.n_states++;
                g.n_clones++;
                Rec CL = Q;
                CL.w[R_LEN] = p_len + 1;
                CL.w[R_OHEAD] = CL.w[R_OTAIL] = -1;
                rec_store(d.recs, clone, CL);
                g.n_edges += rec_free_inline(Q);
                for (uint32_t e = (uint32_t)Q.w[R_OHEAD]; e != SAMD_NIL;) {
                    const uint4 se = slot_load<false>(d.slots, e);
                    ovf_insert(clone, (int)se.y, (int)se.z);
                    SC_STAT(SC_ST_OVF_CLONE);
                    g.n_edges++;
                    e = se.w;
                }
                sc_prefetch_rec(d.recs, clone);             // a fresh record is not in L1 (stores do not allocate)
                // redirect p's suffix chain from q to the clone
                int rp = p, rl = p_link;
                Probe cp = pp;
                while (true) {
                    if (cp.k >= 0) d.recs[(size_t)rp * SAMD_REC + R_TGT + cp.k] = clone;
                    else d.slots[cp.slot].z = (uint32_t)clone;
                    rp = rl;
                    if (rp == -1) break;
                    const Rec R = load(rp);
                    cp = rec_probe<false>(R, d.slots, d.bmask, rp, tok);
                    if (!(cp.found && cp.target == q)) break;
                    rl = R.w[R_LINK];
                }
                d.recs[(size_t)q * SAMD_REC + R_LINK] = clone;
                Q.w[R_LINK] = clone;
                link_cur = clone;
                g.ll_twin = q;
                g.ll_len = p_len + 1;
                g.ll_link = CL.w[R_LINK];
            }
            if (on_edge && q == new_cur) {                  // the cursor moved to q: its record is the next token's
                X = Q;
                x_state = q;
            }
        }
        d.recs[(size_t)cur * SAMD_REC + R_LINK] = link_cur;
        g.last = cur;
        g.last_link = link_cur;
        g.cur = new_cur;
        g.cur_len = new_len;
    }

    // transfer_cur_state only (dyn_sam.py:90-92)
    SAMD_HD void transfer_one(int tok) {
        int x = g.cur, len = g.cur_len;
        if (x_state != x) X = load(x);
        bool first = true;
        while (true) {
            const Probe pr = rec_probe<false>(X, d.slots, d.bmask, x, tok);
            ++g.hops;
            if (!first) len = X.w[R_LEN];
            if (pr.found) {
                g.cur = pr.target;
                g.cur_len = len + 1;
                x_state = -1;
                return;
            }
            if (x == 0) {
                g.cur = 0;
                g.cur_len = 0;
                x_state = 0;
                return;
            }
            x = X.w[R_LINK];
            X = load(x);
            first = false;
        }
    }

    // DynSAM.lookup (dyn_sam.py:94-97): non-mutating peek from the cursor
    SAMD_HD void lookup(int tok, int &index, int &length, int &probes) {
        int x = g.cur, len = g.cur_len;
        Rec Y = (x_state == x) ? X : load(x);
        bool first = true;
        while (true) {
            const Probe pr = rec_probe<false>(Y, d.slots, d.bmask, x, tok);
            ++probes;
            if (!first) len = Y.w[R_LEN];
            if (pr.found) {
                index = pr.target;
                length = len + 1;
                return;
            }
            if (x == 0) {
                index = 0;
                length = 0;
                return;
            }
            x = Y.w[R_LINK];
            Y = load(x);
            first = false;
        }
    }

    // to_anc (dyn_sam.py:99-105) + the draft's anchor: returns min_endpos of the state the draft is read after
    SAMD_HD int anchor_samd(int index, int n_predicts) {
        if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = index;
        int4 h = rec_head<false>(d.recs, index);
        if (index != 0) {
            while (h.x != 0 && n_predicts > g.n - h.z) {
                index = h.x;
                if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = index;
                h = rec_head<false>(d.recs, index);
            }
        }
        return h.z;
    }
};

// ---- read-only cursor walk over the static automaton (static_sam.py:102-109) --------------------------------
template <bool kRO>
SAMD_HD void sc_transfer(const int32_t *recs, const uint4 *slots, uint32_t bmask, int &index, int &length, int tok, int &hops) {
    int x = index, len = length;
    Rec Y = rec_load<kRO>(recs, x);
    bool first = true;
    while (true) {
        const Probe pr = rec_probe<kRO>(Y, slots, bmask, x, tok);
        ++hops;
        if (!first) len = Y.w[R_LEN];
        if (pr.found) {
            index = pr.target;
            length = len + 1;
            return;
        }
        if (x == 0) {
            index = 0;
            length = 0;
            return;
        }
        x = Y.w[R_LINK];
        Y = rec_load<kRO>(recs, x);
        first = false;
    }
}
