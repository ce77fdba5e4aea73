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

// (state, tok) in the overflow table: found -> slot / target; else *free_slot = first free slot of its probe
// sequence (slot by slot from the first slot of the hashed bucket, wrapping: the same order as "the bucket's slots,
// then the next bucket").  Nothing is ever deleted, so a probe sequence is a run of used slots followed by a free
// one, and at the table's load factor (a few percent: only hubs and the root overflow) the first slot decides.
// `max_slots` bounds the search of a reader that races with the writer (scouts); 0 = unbounded.
template <bool kRO>
SAMD_HD bool ovf_find(const uint4 *slots, uint32_t bmask, uint32_t state, uint32_t tok, Probe &r, uint32_t *free_slot,
                      int max_slots = 0) {
    const uint32_t smask = (bmask + 1u) * SAMD_BUCKET - 1u;
    uint32_t i = (samd_hash(state, tok) & bmask) * SAMD_BUCKET;
#pragma unroll 1
    for (int n = 0; max_slots == 0 || n < max_slots; ++n) {
        const uint4 s = slot_load<kRO>(slots, i);
        if (s.x == state && s.y == tok) {
            r.found = true;
            r.slot = i;
            r.target = (int)s.z;
            return true;
        }
        if (s.x == SAMD_EMPTY) {
            if (free_slot) *free_slot = i;
            return false;
        }
        i = (i + 1u) & smask;
    }
    return false;
}

// inline index of the edge on `tok` (which the record is known to have inline)
SAMD_HD int rec_inline_index(const Rec &X, int tok) {
    int kk = 0;
#pragma unroll
    for (int i = 0; i < SAMD_INLINE; ++i) kk |= X.w[R_TOK + i] == tok ? i : 0;
    return kk;
}

// Transition probe.  A state has at most one edge per token, so at most one of the five compares is true: the target is
// OR-ed together from five independent selects (a shallow tree instead of a chain of dependent ones).  r.k is only a
// flag here (0 = inline hit, -1 = not inline): whoever needs the edge's position asks rec_inline_index afterwards - one
// probe per chain needs it, every stop of every chain is probed.
template <bool kRO>
SAMD_HD Probe rec_probe(const Rec &X, const uint4 *slots, uint32_t bmask, int state, int tok, int max_slots = 0) {
    Probe r;
    int tgt = 0;
    bool any = false;
#pragma unroll
    for (int i = 0; i < SAMD_INLINE; ++i) {
        const bool e = X.w[R_TOK + i] == tok;
        tgt |= e ? X.w[R_TGT + i] : 0;
        any |= e;
    }
    r.target = tgt;
    r.k = any ? 0 : -1;
    r.slot = SAMD_NIL;
    r.found = any;
    if (!any && (uint32_t)X.w[R_OHEAD] != SAMD_NIL) ovf_find<kRO>(slots, bmask, (uint32_t)state, (uint32_t)tok, r, nullptr, max_slots);
    return r;
}

// ---- the builder -----------------------------------------------------------------------------------------
// Arena invariants the builder relies on (dyn_reset_kernel / samd_dyn_create establish them, both kernel variants keep
// them): every record beyond n_states holds the EMPTY TEMPLATE {link -1, length 0, min_endpos 0, no edges, no overflow
// list, aux 0}, so a new state costs three 4-byte stores; and a record's spare word R_AUX counts its inline edges, so
// the free inline slot of a state is known from its record without comparing anything.
struct ScDyn {
    int32_t *recs;
    uint4   *slots;
    int32_t *text;
    uint32_t bmask;
    int      max_tokens;
};

// ll_*: when the newest append split a state, last_link is that clone and `ll_twin` the state it was copied from -
// their inline edges are identical, so a probe of the twin (the cursor's state) answers for the clone as well.
struct ScRegs {
    int n_states, last, last_link, n, cur, cur_len, n_edges, n_clones, hops;
    int ll_twin, ll_len, ll_link;
    int max_chain;
};

struct ScCounters {            // profiling build: per-request trace of the record reads (floor analysis), see tools/
    int32_t *trace;
    int      cap, n;
};

#ifdef SAMD_SCALAR_STATS          // host test build only: which paths of extend_one ran
#define SC_STAT(x) (++stats[x])
#else
#define SC_STAT(x) ((void)0)
#endif
enum { SC_ST_ALIGNED = 0, SC_ST_TWIN = 1, SC_ST_GENERIC = 2, SC_ST_LONGCHAIN = 3, SC_ST_OVF_INSERT = 4, SC_ST_OVF_CLONE = 5,
       SC_ST_CARRIED = 6, SC_ST_REDIR = 7, SC_ST_N = 8 };

// cycle counter + a consumer of a load's result, for the profiling build (the clock is read after the data arrived)
SAMD_HD long long sc_clock() {
#if defined(__CUDA_ARCH__)
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
#else
    return 0;
#endif
}
SAMD_HD void sc_consume(int a, int b) {
#if defined(__CUDA_ARCH__)
    asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; }" ::"r"(a), "r"(b) : "memory");
#else
    (void)a; (void)b;
#endif
}
enum { SC_PF_LOAD_CYC = 0, SC_PF_L1 = 1, SC_PF_L2 = 2, SC_PF_DRAM = 3, SC_PF_SLOW = 4, SC_PF_OVF_CYC = 5, SC_PF_OVF_N = 6, SC_PF_WALK_CYC = 7,
       SC_PF_REDIR_CYC = 8, SC_PF_REDIR_N = 9, SC_PF_N = 10 };

template <bool kProf>
struct ScBuilderT {
    long long pf[kProf ? SC_PF_N : 1];   // profiling build: cycles waiting for record loads, loads by latency class, overflow probes
#ifdef SAMD_SCALAR_STATS
    long long stats[SC_ST_N];
#endif
    ScDyn   d;
    ScRegs  g;
    Rec     X;                 // the cursor's record when x_state == g.cur (carried from token to token)
    int     x_state;
    ScCounters tr;

    SAMD_HD Rec load(int state) {
        if constexpr (kProf) {
            if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = state;
            const long long t0 = sc_clock();
            const Rec R = rec_load<false>(d.recs, state);
            sc_consume(R.w[0], R.w[15]);
            const long long dt = sc_clock() - t0;
            pf[SC_PF_LOAD_CYC] += dt;
            pf[SC_PF_L1] += dt < 120;
            pf[SC_PF_L2] += dt >= 120 && dt < 500;
            pf[SC_PF_DRAM] += dt >= 500 && dt < 1100;
            pf[SC_PF_SLOW] += dt >= 1100;
            return R;
        } else {
            return rec_load<false>(d.recs, state);
        }
    }
    SAMD_HD Probe probe(const Rec &Y, int state, int tok) {
        if constexpr (kProf) {
            if ((uint32_t)Y.w[R_OHEAD] != SAMD_NIL) {
                const long long t0 = sc_clock();
                const Probe r = rec_probe<false>(Y, d.slots, d.bmask, state, tok);
                sc_consume(r.target, r.k);
                pf[SC_PF_OVF_CYC] += sc_clock() - t0;
                pf[SC_PF_OVF_N] += 1;
                return r;
            }
        }
        return rec_probe<false>(Y, d.slots, d.bmask, state, tok);
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

    // give `state` (which has `n_inline` inline edges and none on tok) the edge tok -> target
    SAMD_HD void insert_edge(int state, int n_inline, int tok, int target) {
        if (n_inline < SAMD_INLINE) {
            int32_t *rec = d.recs + (size_t)state * SAMD_REC;
            rec[R_TOK + n_inline] = tok;
            rec[R_TGT + n_inline] = target;
            rec[R_AUX] = n_inline + 1;
        } else {
            SC_STAT(SC_ST_OVF_INSERT);
            ovf_insert(state, tok, target);
        }
    }

    // transfer_cur_state + add_state for one token (dyn_sam.py:84-88: match first, then append).
    //
    // The cursor's fallback walk for the token (follow suffix links until a state has an edge on it) visits exactly
    // the states add_state gives that edge to: add_state walks the suffix chain of `last` from link(last), and the
    // cursor IS link(last) - or, right after a split, the pre-clone state one stop in front of it.  So the two walks
    // are ONE: a stop that lacks the edge gets it on the spot, from the record the probe already holds.
    SAMD_HD void extend_one(int tok) {
        long long t_walk0 = 0;
        if constexpr (kProf) t_walk0 = sc_clock();
        int x = g.cur, len = g.cur_len;
        if (x_state != x) X = load(x);
        else SC_STAT(SC_ST_CARRIED);
        const int cur = g.n_states;                          // the state this append creates
        const int last = g.last;
        const int p0 = last != 0 ? g.last_link : 0;          // first state of add_state's chain that has to be read
        bool aligned = (x == p0);                            // the cursor's walk is (from here on) add_state's walk
        bool first = true, on_edge = false;
        int n_chain = 0, n_ins = 0;
        Probe pr;
        while (true) {
            pr = probe(X, x, tok);
            ++n_chain;
            if (!first) len = X.w[R_LEN];                    // length = states[index].length after a link hop
            if (pr.found) {
                on_edge = true;
                break;
            }
            if (x == 0) {
                if (aligned) {
                    insert_edge(0, X.w[R_AUX], tok, cur);
                    ++n_ins;
                }
                break;
            }
            const int nx = X.w[R_LINK], n_inline = X.w[R_AUX];
            X = load(nx);                                    // requested before this stop's stores are issued
            if (aligned) {
                insert_edge(x, n_inline, tok, cur);
                ++n_ins;
            } else if (first && nx == p0) {
                aligned = true;                              // x was the pre-clone state in front of the chain
            }
            x = nx;
            first = false;
        }
        if constexpr (kProf) pf[SC_PF_WALK_CYC] += sc_clock() - t_walk0;      // the cursor's walk (with the fused inserts)
        g.hops += n_chain;
        if (n_chain > g.max_chain) g.max_chain = n_chain;
        const int new_cur = on_edge ? pr.target : 0;
        const int new_len = on_edge ? len + 1 : 0;

        // ---- add_state (dyn_sam.py:41-67) ----
        g.n += 1;
        g.n_states = cur + 1;
        if (last != 0) {
            // `last` was created by the previous append and has no out-edge yet: its first inline edge needs no read
            int32_t *rl = d.recs + (size_t)last * SAMD_REC;
            rl[R_TOK] = tok;
            rl[R_TGT] = cur;
            rl[R_AUX] = 1;
            ++n_ins;
        }
        g.n_edges += n_ins;
        int p = -1, p_len = 0, p_link = -1;                  // the chain state that has the edge (-1: none, link -> root)
        Probe pp = pr;
        if (aligned) {
            SC_STAT(SC_ST_ALIGNED);
            if (on_edge) {
                p = x;
                p_len = X.w[R_LEN];
                p_link = X.w[R_LINK];
                if (pp.k >= 0) pp.k = rec_inline_index(X, tok);
            }
        } else if (on_edge && n_chain == 1 && last != 0 && g.ll_twin == x && pr.k >= 0) {
            // last_link is the clone the previous append made of the cursor's state: same inline edges, so the probe
            // of the cursor's record stands for it
            SC_STAT(SC_ST_TWIN);
            p = p0;
            p_len = g.ll_len;
            p_link = g.ll_link;
            pp.k = rec_inline_index(X, tok);
        } else {
            // the cursor is somewhere else (DynSAM.transfer_tokens moved it, or a kernel boundary dropped the twin
            // hint): add_state's own walk
            SC_STAT(SC_ST_GENERIC);
            p = p0;
            while (p != -1) {
                const Rec P = load(p);
                const Probe q = probe(P, p, tok);
                if (!q.found) {
                    insert_edge(p, P.w[R_AUX], tok, cur);
                    g.n_edges++;
                    p = P.w[R_LINK];
                    continue;
                }
                p_len = P.w[R_LEN];
                p_link = P.w[R_LINK];
                pp = q;
                if (pp.k >= 0) pp.k = rec_inline_index(P, tok);
                break;
            }
        }
        const int q = pp.target;
        if (p != -1) X = load(q);                            // after every store above (q may be `last` or a chain state)
        {
            // the new state: the empty template is already in place; these stores ride in the shadow of that load
            int32_t *rc = d.recs + (size_t)cur * SAMD_REC;
            rc[R_LEN] = g.n;
            rc[R_END] = g.n;
            d.text[g.n] = tok;
        }
        int link_cur = 0;
        g.ll_twin = -1;
        x_state = -1;
        if (p != -1) {
            if (p_len + 1 == X.w[R_LEN]) {
                link_cur = q;
            } else {
                // clone-on-split: q's record with length len(p)+1; overflow edges re-inserted oldest first
                const int clone = g.n_states++;
                g.n_clones++;
                const int q_len = X.w[R_LEN], q_oh = X.w[R_OHEAD], q_ot = X.w[R_OTAIL], q_link = X.w[R_LINK];
                X.w[R_LEN] = p_len + 1;
                X.w[R_OHEAD] = X.w[R_OTAIL] = -1;
                rec_store(d.recs, clone, X);
                X.w[R_LEN] = q_len;
                X.w[R_OHEAD] = q_oh;
                X.w[R_OTAIL] = q_ot;
                g.n_edges += X.w[R_AUX];
                for (uint32_t e = (uint32_t)q_oh; e != SAMD_NIL;) {
                    const uint4 se = slot_load<false>(d.slots, e);
                    ovf_insert(clone, (int)se.y, (int)se.z);
                    SC_STAT(SC_ST_OVF_CLONE);
                    g.n_edges++;
                    e = se.w;
                }
                long long t_red0 = 0;
                if constexpr (kProf) t_red0 = sc_clock();
                // redirect p's suffix chain from q to the clone
                int rp = p, rl = p_link;
                Probe cp = pp;
                while (true) {
                    if (cp.k >= 0) d.recs[(size_t)rp * SAMD_REC + R_TGT + cp.k] = clone;
                    else d.slots[cp.slot].z = (uint32_t)clone;
                    rp = rl;
                    if (rp == -1) break;
                    const Rec R = load(rp);
                    if constexpr (kProf) pf[SC_PF_REDIR_N] += 1;
                    cp = probe(R, rp, tok);
                    if (!(cp.found && cp.target == q)) break;
                    SC_STAT(SC_ST_REDIR);
                    if (cp.k >= 0) cp.k = rec_inline_index(R, tok);
                    rl = R.w[R_LINK];
                }
                if constexpr (kProf) pf[SC_PF_REDIR_CYC] += sc_clock() - t_red0;
                d.recs[(size_t)q * SAMD_REC + R_LINK] = clone;
                X.w[R_LINK] = clone;
                sc_prefetch_rec(d.recs, clone);             // a record that was only written is not in L1 (stores do not allocate)
                link_cur = clone;
                g.ll_twin = q;
                g.ll_len = p_len + 1;
                g.ll_link = q_link;
            }
            if (on_edge && q == new_cur) x_state = q;       // the cursor moved to q: X is the next token's cursor record
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
            const Probe pr = probe(X, x, tok);
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
            const Probe pr = probe(Y, x, tok);
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
        if constexpr (kProf)
            if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = index;
        int4 h = rec_head<false>(d.recs, index);
        if (index != 0) {
            while (h.x != 0 && n_predicts > g.n - h.z) {
                index = h.x;
                if constexpr (kProf)
                    if (tr.trace && tr.n < tr.cap) tr.trace[tr.n++] = index;
                h = rec_head<false>(d.recs, index);
            }
        }
        return h.z;
    }
};
typedef ScBuilderT<false> ScBuilder;

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
