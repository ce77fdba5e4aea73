/*
 * sam_oracle.c - plain-C restatement of the reference's suffix-automaton path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/samd_oracle.py for the rules): built into
 * oracle/_build/libsam_oracle.so by oracle/Makefile, loaded by tests/ and by the cpu_baseline legs of
 * bench.py.  It exists so that parity can be checked at BASELINE.json's full sizes (1024 x 8k-token
 * requests, multi-million-token corpora), where pure-Python loops are too slow, and so that a
 * compiled CPU baseline can be reported next to the Python port.
 *
 * Parity status: pinned through tests/test_oracle_golden.py::test_c_oracle_* against the fixtures the
 * reference's own classes produced (tests/golden/), and against oracle/samd_oracle.py.
 *
 * Deliberately different from the product's layout: one open-addressing table of 64-bit keys
 * (state << 32 | token) -> target per automaton and a newest-first intrusive edge list per state.
 *
 * Reference lines restated:
 *   so_append      samd/sam/dyn_sam.py:41-67     (add_state, clone-on-split)
 *   so_step        samd/sam/dyn_sam.py:69-78     (transfer_state)
 *   so_extend      samd/sam/dyn_sam.py:84-88     (add_tokens: match first, then append)
 *   so_advance     samd/sam/dyn_sam.py:90-92     (transfer_tokens)
 *   so_peek        samd/sam/dyn_sam.py:94-97     (lookup)
 *   so_draft_samd  samd/sam/dyn_sam.py:99-113    (to_anc + gen_draft; anchor=0 -> static_sam.py:119-125)
 *   so_draft_so    samd_sam_only/sam/dyn_sam.py:116-121
 *   so_select_samd samd/draft.py:52-63
 *   so_build_docs  samd/sam/static_sam.py:32-46
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t *link, *len, *end, *head;     /* per state */
    uint64_t *key;                        /* hash keys, 0 = free (state+1 in the high word) */
    int32_t *val, *nxt;                   /* target, next edge of the same state */
    int32_t *text;                        /* 1-based */
    int64_t cap_states, cap_slots, mask;
    int32_t n_states, last, n, cur, cur_len;
    int64_t n_edges, n_clones;
} sam_t;

static uint64_t mix(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

sam_t *so_new(int64_t max_tokens) {
    sam_t *s = (sam_t *)calloc(1, sizeof(sam_t));
    s->cap_states = 2 * max_tokens + 2;
    int64_t slots = 64;
    while (slots < 6 * (max_tokens + 1)) slots <<= 1;
    s->cap_slots = slots;
    s->mask = slots - 1;
    s->link = (int32_t *)malloc(s->cap_states * 4);
    s->len = (int32_t *)malloc(s->cap_states * 4);
    s->end = (int32_t *)malloc(s->cap_states * 4);
    s->head = (int32_t *)malloc(s->cap_states * 4);
    s->key = (uint64_t *)calloc(slots, 8);
    s->val = (int32_t *)malloc(slots * 4);
    s->nxt = (int32_t *)malloc(slots * 4);
    s->text = (int32_t *)malloc((max_tokens + 2) * 4);
    s->link[0] = -1;
    s->len[0] = 0;
    s->end[0] = 0;
    s->head[0] = -1;
    s->text[0] = -1;
    s->n_states = 1;
    return s;
}

void so_free(sam_t *s) {
    if (!s) return;
    free(s->link); free(s->len); free(s->end); free(s->head);
    free(s->key); free(s->val); free(s->nxt); free(s->text);
    free(s);
}

static inline uint64_t mk(int32_t state, int32_t tok) { return ((uint64_t)(uint32_t)(state + 1) << 32) | (uint32_t)tok; }

/* slot of (state, tok) or -1 */
static inline int64_t find(const sam_t *s, int32_t state, int32_t tok) {
    const uint64_t k = mk(state, tok);
    int64_t i = (int64_t)(mix(k) & (uint64_t)s->mask);
    while (s->key[i]) {
        if (s->key[i] == k) return i;
        i = (i + 1) & s->mask;
    }
    return -1;
}

static inline void put(sam_t *s, int32_t state, int32_t tok, int32_t target) {
    const uint64_t k = mk(state, tok);
    int64_t i = (int64_t)(mix(k) & (uint64_t)s->mask);
    while (s->key[i]) i = (i + 1) & s->mask;
    s->key[i] = k;
    s->val[i] = target;
    s->nxt[i] = s->head[state];
    s->head[state] = (int32_t)i;
    s->n_edges++;
}

static inline int32_t new_state(sam_t *s, int32_t link, int32_t len, int32_t end) {
    const int32_t v = s->n_states++;
    s->link[v] = link;
    s->len[v] = len;
    s->end[v] = end;
    s->head[v] = -1;
    return v;
}

void so_append(sam_t *s, int32_t tok) {
    s->n += 1;
    const int32_t cur = new_state(s, -1, s->n, s->n);
    int32_t p = s->last;
    int64_t e = -1;
    while (p != -1 && (e = find(s, p, tok)) < 0) {
        put(s, p, tok, cur);
        p = s->link[p];
    }
    if (p == -1) {
        s->link[cur] = 0;
    } else {
        const int32_t q = s->val[e];
        if (s->len[p] + 1 == s->len[q]) {
            s->link[cur] = q;
        } else {
            const int32_t clone = new_state(s, s->link[q], s->len[p] + 1, s->end[q]);
            s->n_clones++;
            for (int32_t i = s->head[q]; i != -1; i = s->nxt[i]) put(s, clone, (int32_t)(uint32_t)s->key[i], s->val[i]);
            while (p != -1 && (e = find(s, p, tok)) >= 0 && s->val[e] == q) {
                s->val[e] = clone;
                p = s->link[p];
            }
            s->link[q] = clone;
            s->link[cur] = clone;
        }
    }
    s->last = cur;
    s->text[s->n] = tok;
}

void so_step(const sam_t *s, int32_t *state, int32_t *matched, int32_t tok) {
    int32_t v = *state, m = *matched;
    int64_t e;
    while (v != 0 && (e = find(s, v, tok)) < 0) {
        v = s->link[v];
        m = s->len[v];
    }
    e = find(s, v, tok);
    if (e >= 0) {
        *state = s->val[e];
        *matched = m + 1;
    } else {
        *state = 0;
        *matched = 0;
    }
}

void so_extend(sam_t *s, const int32_t *tokens, int64_t k) {
    for (int64_t i = 0; i < k; ++i) {
        so_step(s, &s->cur, &s->cur_len, tokens[i]);
        so_append(s, tokens[i]);
    }
}

void so_advance(sam_t *s, const int32_t *tokens, int64_t k) {
    for (int64_t i = 0; i < k; ++i) so_step(s, &s->cur, &s->cur_len, tokens[i]);
}

void so_reset_cursor(sam_t *s) { s->cur = 0; s->cur_len = 0; }

void so_peek(const sam_t *s, int32_t tok, int32_t *state, int32_t *matched) {
    *state = s->cur;
    *matched = s->cur_len;
    so_step(s, state, matched, tok);
}

/* [start] + text[e+1 : e+n], zero padded to n; anchor != 0 applies the dynamic automaton's to_anc walk */
void so_draft_samd(const sam_t *s, int32_t state, int32_t start_tok, int32_t n, int32_t anchor, int32_t *out) {
    if (anchor && state != 0) {
        while (s->link[state] != 0 && n > s->n - s->end[state]) state = s->link[state];
    }
    const int32_t e = s->end[state];
    out[0] = start_tok;
    for (int32_t j = 1; j < n; ++j) out[j] = (e + j <= s->n) ? s->text[e + j] : 0;
}

int32_t so_draft_so(const sam_t *s, int32_t state, int32_t matched, int32_t start_tok, int32_t max_predicts, double alpha,
                    int32_t *out) {
    int32_t n = 1 + (int32_t)((double)matched * alpha);
    if (n > max_predicts) n = max_predicts;
    const int32_t e = s->end[state];
    int32_t k = 0;
    out[k++] = start_tok;
    for (int32_t j = 1; j < n && e + j <= s->n; ++j) out[k++] = s->text[e + j];
    return k;
}

/* DraftModel.lookup of `samd`: returns 0 dyn sequence, 1 static sequence, 2 tree fallback */
int32_t so_select_samd(const sam_t *dyn, sam_t *stat, int32_t start_tok, int32_t n, int32_t len_bias, int32_t len_threshold,
                       int32_t *out, int32_t *info) {
    int32_t si, sl, ti = 0, tl = 0;
    so_peek(dyn, start_tok, &si, &sl);
    if (stat) so_peek(stat, start_tok, &ti, &tl);
    if (info) { info[0] = si; info[1] = sl; info[2] = ti; info[3] = tl; }
    const int32_t tb = tl - len_bias;
    if ((sl > tb ? sl : tb) >= len_threshold) {
        if (sl >= tb) {
            so_draft_samd(dyn, si, start_tok, n, 1, out);
            return 0;
        }
        so_draft_samd(stat, ti, start_tok, n, 0, out);
        return 1;
    }
    return 2;
}

sam_t *so_build_docs(const int32_t *flat, const int64_t *offs, int64_t n_docs, int32_t eos) {
    int64_t total = 0;
    for (int64_t d = 0; d < n_docs; ++d) total += offs[d + 1] - offs[d] + 1;
    sam_t *s = so_new(total);
    for (int64_t d = 0; d < n_docs; ++d) {
        so_extend(s, flat + offs[d], offs[d + 1] - offs[d]);
        if (flat[offs[d + 1] - 1] != eos) so_extend(s, &eos, 1);
    }
    so_reset_cursor(s);
    return s;
}

void so_info(const sam_t *s, int64_t *out) {
    out[0] = s->n_states; out[1] = s->n; out[2] = s->n_edges; out[3] = s->n_clones;
    out[4] = s->cur; out[5] = s->cur_len; out[6] = s->last;
}

void so_export(const sam_t *s, int32_t *link, int32_t *len, int32_t *end) {
    memcpy(link, s->link, (size_t)s->n_states * 4);
    memcpy(len, s->len, (size_t)s->n_states * 4);
    memcpy(end, s->end, (size_t)s->n_states * 4);
}

/* CPU baseline driver: the c2 step loop for a block of requests (bench.py).  tokens [S][R][8], counts [S][R],
 * start [S][R]; returns a checksum of all drafts so the work cannot be optimised away. */
int64_t so_run_steps(sam_t **sams, int64_t R, const int32_t *tokens, const int32_t *counts, const int32_t *start, int64_t s_lo,
                     int64_t s_hi, int32_t n_predicts, int32_t len_bias, int32_t len_threshold) {
    int32_t draft[256];
    int64_t sum = 0;
    for (int64_t s = s_lo; s < s_hi; ++s)
        for (int64_t r = 0; r < R; ++r) {
            so_extend(sams[r], tokens + (s * R + r) * 8, counts[s * R + r]);
            if (so_select_samd(sams[r], NULL, start[s * R + r], n_predicts, len_bias, len_threshold, draft, NULL) != 2)
                for (int32_t j = 0; j < n_predicts; ++j) sum += draft[j];
        }
    return sum;
}

/* Many independent query cursors over ONE read-only automaton (the static SAM's use, static_sam.py:102-125): cur[i] =
 * {state, matched}.  so_batch_advance = StaticSAM.transfer_tokens for every cursor (counts NULL = stride tokens each);
 * so_batch_lookup = StaticSAM.lookup + gen_draft (no to_anc) with each cursor's start token, cursors unchanged. */
void so_batch_advance(const sam_t *s, int32_t *cur, const int32_t *tokens, int64_t stride, const int32_t *counts, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t k = counts ? counts[i] : stride;
        for (int64_t j = 0; j < k; ++j) so_step(s, &cur[2 * i], &cur[2 * i + 1], tokens[i * stride + j]);
    }
}

void so_batch_lookup(const sam_t *s, const int32_t *cur, const int32_t *start, int64_t n, int32_t n_predicts, int32_t *out_state,
                     int32_t *out_len, int32_t *out_draft) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t st = cur[2 * i], ln = cur[2 * i + 1];
        so_step(s, &st, &ln, start[i]);
        out_state[i] = st;
        out_len[i] = ln;
        if (out_draft) so_draft_samd(s, st, start[i], n_predicts, 0, out_draft + i * n_predicts);
    }
}

int32_t so_min_endpos(const sam_t *s, int32_t state) { return s->end[state]; }
