// TEST INFRASTRUCTURE ONLY: compiles the product's one-thread step logic (sam-decoding_b200/csrc/sam_scalar.cuh) for
// the host, so that its append / clone / redirect / chain-insert paths can be fuzzed against the oracle on a machine
// without a GPU (tests/test_scalar_core_cpu.py).  Never loaded by the product; the product path is the CUDA kernel.
#define SAMD_SCALAR_STATS 1
#include "../../sam-decoding_b200/csrc/sam_scalar.cuh"

#include <stdlib.h>
#include <vector>

void samd_set_error(const char *, ...) {}
void samd_count_launch(int) {}

struct Emul {
    ScBuilder b;
    std::vector<int32_t> recs, text;
    std::vector<uint4> slots;
};

extern "C" {

Emul *emul_new(int max_tokens) {
    Emul *e = new Emul();
    const uint32_t s_cap = 2u * (uint32_t)max_tokens + 2u;
    const uint32_t h_cap = (uint32_t)samd_table_slots((uint64_t)max_tokens);
    e->recs.assign((size_t)s_cap * SAMD_REC, 0);
    e->slots.assign(h_cap, make_uint4(SAMD_EMPTY, SAMD_EMPTY, SAMD_EMPTY, SAMD_EMPTY));
    e->text.assign((size_t)max_tokens + 4, 0);
    for (uint32_t v = 0; v < s_cap; ++v) samd_init_rec(e->recs.data() + (size_t)v * SAMD_REC, -1, 0, 0);   // the arena invariant: empty templates
    e->text[0] = -1;
    ScBuilder &b = e->b;
    b.d.recs = e->recs.data();
    b.d.slots = e->slots.data();
    b.d.text = e->text.data();
    b.d.bmask = h_cap / SAMD_BUCKET - 1u;
    b.d.max_tokens = max_tokens;
    b.g = ScRegs{1, 0, -1, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0};
    b.x_state = -1;
    for (int i = 0; i < SC_ST_N; ++i) b.stats[i] = 0;
    b.tr.trace = nullptr;
    b.tr.cap = b.tr.n = 0;
    return e;
}

void emul_free(Emul *e) { delete e; }

// a kernel boundary: registers are gone, only what the meta block keeps survives
void emul_boundary(Emul *e) { e->b.x_state = -1; }

int emul_extend(Emul *e, const int32_t *tok, int k) {
    for (int i = 0; i < k; ++i) {
        if (e->b.g.n >= e->b.d.max_tokens) return 1;
        e->b.extend_one(tok[i]);
    }
    return 0;
}

void emul_transfer(Emul *e, const int32_t *tok, int k) {
    for (int i = 0; i < k; ++i) e->b.transfer_one(tok[i]);
}

void emul_lookup(Emul *e, int tok, int32_t *index, int32_t *length) {
    int probes = 0, i = 0, l = 0;
    e->b.lookup(tok, i, l, probes);
    *index = i;
    *length = l;
}

// samd flavour: to_anc + [start] + text[e+1 : e+n], zero padded
void emul_draft_samd(Emul *e, int index, int start_tok, int n, int32_t *out) {
    const int endpos = e->b.anchor_samd(index, n);
    out[0] = start_tok;
    for (int j = 1; j < n; ++j) out[j] = (endpos + j <= e->b.g.n) ? e->text[endpos + j] : 0;
}

void emul_info(Emul *e, int64_t *out) {
    const ScRegs &g = e->b.g;
    out[0] = g.n_states; out[1] = g.n; out[2] = g.n_edges; out[3] = g.n_clones;
    out[4] = g.cur; out[5] = g.cur_len; out[6] = g.last; out[7] = g.last_link;
}

// which paths of extend_one ran: {aligned, twin, generic, (unused), overflow inserts, overflow edges cloned, carried record}
void emul_stats(Emul *e, int64_t *out) {
    for (int i = 0; i < SC_ST_N; ++i) out[i] = e->b.stats[i];
}

void emul_export(Emul *e, int32_t *link, int32_t *len, int32_t *end) {
    for (int v = 0; v < e->b.g.n_states; ++v) {
        link[v] = e->recs[(size_t)v * SAMD_REC + R_LINK];
        len[v] = e->recs[(size_t)v * SAMD_REC + R_LEN];
        end[v] = e->recs[(size_t)v * SAMD_REC + R_END];
    }
}

// edges as (state, token, target), per state in insertion order; returns the count
int64_t emul_export_edges(Emul *e, int32_t *out, int64_t cap) {
    int64_t k = 0;
    for (int v = 0; v < e->b.g.n_states; ++v) {
        const int32_t *rec = e->recs.data() + (size_t)v * SAMD_REC;
        for (int i = 0; i < SAMD_INLINE && (uint32_t)rec[R_TOK + i] != SAMD_EMPTY; ++i, ++k)
            if (k < cap) { out[3 * k] = v; out[3 * k + 1] = rec[R_TOK + i]; out[3 * k + 2] = rec[R_TGT + i]; }
        for (uint32_t s = (uint32_t)rec[R_OHEAD]; s != SAMD_NIL; s = e->slots[s].w, ++k)
            if (k < cap) { out[3 * k] = v; out[3 * k + 1] = (int32_t)e->slots[s].y; out[3 * k + 2] = (int32_t)e->slots[s].z; }
    }
    return k;
}
}
