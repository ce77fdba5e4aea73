"""CPU oracle for the SAM-Decoding draft-retrieval + verification hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain-Python / numpy restatement of the
reference algorithm (hyx1999/SAM-Decoding).  It may be imported by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`, and only as the checker / timed CPU baseline - never by the product
path under `sam-decoding_b200/`.

Parity status: PINNED.  The reference ships no golden vectors of its own
(SURVEY.md §4, §8c), so the pin is differential: `oracle/gen_golden.py` runs the
reference's own classes (imported from /root/reference in the build container) on
seeded inputs and commits their outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those fixtures.

Every function cites the reference file:line it restates (paths relative to the
reference checkout).  Layout differs from the reference on purpose (struct of
arrays, free functions): it is a restatement, not a copy.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

ROOT = 0


# --------------------------------------------------------------------------------------
# Suffix automaton core (shared by the dynamic and the static automaton)
# --------------------------------------------------------------------------------------
class Automaton:
    """Online suffix automaton, struct-of-arrays.

    Restates samd/sam/dyn_sam.py:8-97 (== samd/sam/static_sam.py:8-109,
    samd_sam_only/sam/dyn_sam.py:11-114, samd_sam_only/sam/static_sam.py:22-124).
    State 0 is the root (link -1, length 0, first_end 0).  Text positions are
    1-based; text[0] is a -1 sentinel (dyn_sam.py:20).
    """

    def __init__(self, count_occurrences: bool = False, keep_text: bool = True):
        self.count_occurrences = count_occurrences   # samd_sam_only static flavour
        self.keep_text = keep_text                   # sam_only static keeps no token array
        self.clear()

    def clear(self):
        self.link: List[int] = [-1]
        self.length: List[int] = [0]
        self.first_end: List[int] = [0]              # `min_endpos`
        self.occ: List[int] = [0]                    # `cnt_endpos` (count flavour only)
        self.trans: List[Dict[int, int]] = [{}]      # insertion-ordered, like the reference dicts
        self.text: List[int] = [-1]
        self.tail = ROOT                             # `last`
        self.n = 0                                   # `max_length`
        self.cur = ROOT                              # query cursor state
        self.cur_len = 0                             # query cursor match length
        self.n_clones = 0
        self.n_chain_hops = 0

    # -- dyn_sam.py:41-67 ------------------------------------------------------------
    def append(self, tok: int) -> None:
        self.n += 1
        new = len(self.link)
        self.link.append(-1)
        self.length.append(self.n)
        self.first_end.append(self.n)
        self.occ.append(0)
        self.trans.append({})
        p = self.tail
        while p != -1 and tok not in self.trans[p]:
            self.trans[p][tok] = new
            p = self.link[p]
            self.n_chain_hops += 1
        if p == -1:
            self.link[new] = ROOT
        else:
            q = self.trans[p][tok]
            if self.length[p] + 1 == self.length[q]:
                self.link[new] = q
            else:
                # clone-on-split: the copy inherits edges, link, first_end (and count)
                cl = len(self.link)
                self.link.append(self.link[q])
                self.length.append(self.length[p] + 1)
                self.first_end.append(self.first_end[q])
                self.occ.append(self.occ[q])
                self.trans.append(dict(self.trans[q]))
                self.n_clones += 1
                while p != -1 and self.trans[p][tok] == q:
                    self.trans[p][tok] = cl
                    p = self.link[p]
                self.link[q] = cl
                self.link[new] = cl
        self.tail = new
        if self.count_occurrences:
            # samd_sam_only/sam/static_sam.py:93-96
            s = new
            while s != ROOT:
                self.occ[s] += 1
                s = self.link[s]

    # -- dyn_sam.py:69-78 ------------------------------------------------------------
    def step(self, state: int, matched: int, tok: int) -> Tuple[int, int]:
        """One longest-suffix-match transition; pure function of the automaton."""
        while state != ROOT and tok not in self.trans[state]:
            state = self.link[state]
            matched = self.length[state]
        nxt = self.trans[state].get(tok)
        if nxt is None:
            return ROOT, 0
        return nxt, matched + 1

    # -- dyn_sam.py:84-88: match first, then append ------------------------------------
    def extend(self, tokens: Iterable[int]) -> None:
        tokens = [int(t) for t in tokens]
        for t in tokens:
            self.cur, self.cur_len = self.step(self.cur, self.cur_len, t)
            self.append(t)
        if self.keep_text:
            self.text.extend(tokens)

    # -- dyn_sam.py:90-92 ------------------------------------------------------------
    def advance(self, tokens: Iterable[int]) -> None:
        for t in tokens:
            self.cur, self.cur_len = self.step(self.cur, self.cur_len, int(t))

    # -- dyn_sam.py:94-97 ------------------------------------------------------------
    def peek(self, tok: int) -> Tuple[int, int]:
        return self.step(self.cur, self.cur_len, int(tok))

    def reset_cursor(self):          # static_sam.py:28-30
        self.cur, self.cur_len = ROOT, 0

    @property
    def n_states(self) -> int:
        return len(self.link)

    @property
    def n_edges(self) -> int:
        return sum(len(t) for t in self.trans)


# --------------------------------------------------------------------------------------
# Draft extraction, `samd` flavour
# --------------------------------------------------------------------------------------
def dyn_anchor(sam: Automaton, state: int, n_predicts: int) -> int:
    """samd/sam/dyn_sam.py:99-105 (`to_anc`): climb suffix links while the earliest
    occurrence is too close to the end of the text to supply n_predicts tokens; never
    steps onto the root."""
    if state != ROOT:
        room = sam.n - sam.first_end[state]
        while sam.link[state] != ROOT and n_predicts > room:
            state = sam.link[state]
            room = sam.n - sam.first_end[state]
    return state


def _slice_pad(text: List[int], start_tok: int, end: int, n: int) -> List[int]:
    out = [int(start_tok)] + text[end + 1:end + n]
    out.extend([0] * (n - len(out)))
    return out


def dyn_draft_samd(sam: Automaton, state: int, start_tok: int, n_predicts: int) -> List[int]:
    """samd/sam/dyn_sam.py:107-113."""
    state = dyn_anchor(sam, state, n_predicts)
    return _slice_pad(sam.text, start_tok, sam.first_end[state], n_predicts)


def static_draft_samd(sam: Automaton, state: int, start_tok: int, n_predicts: int) -> List[int]:
    """samd/sam/static_sam.py:119-125 (no anchor walk; may run across EOS)."""
    return _slice_pad(sam.text, start_tok, sam.first_end[state], n_predicts)


def select_samd(dyn: Automaton, static: Optional[Automaton], start_tok: int, n_predicts: int,
                len_bias: int, len_threshold: int):
    """samd/draft.py:52-63.  Returns ("sequence", draft, info) or ("tree", None, info) when the
    biased match is below the threshold (the tree-model fallback is out of scope)."""
    si, sl = dyn.peek(start_tok)
    if static is None:                 # NullStaticSAM: root-only automaton
        ti, tl = ROOT, 0
    else:
        ti, tl = static.peek(start_tok)
    info = dict(index_dyn=si, match_dyn=sl, index_static=ti, match_static=tl)
    tl_b = tl - len_bias
    if max(sl, tl_b) >= len_threshold:
        if sl >= tl_b:
            info["source"] = "dyn"
            return "sequence", dyn_draft_samd(dyn, si, start_tok, n_predicts), info
        info["source"] = "static"
        return "sequence", static_draft_samd(static, ti, start_tok, n_predicts), info
    info["source"] = "tree"
    return "tree", None, info


# --------------------------------------------------------------------------------------
# Draft extraction, `samd_sam_only` flavour
# --------------------------------------------------------------------------------------
def draft_budget(match_length: int, max_predicts: int, alpha: float) -> int:
    """samd_sam_only/sam/dyn_sam.py:117: n = min(max_predicts, 1 + int(match * alpha))."""
    return min(max_predicts, 1 + int(match_length * alpha))


def dyn_draft_sam_only(sam: Automaton, state: int, match_length: int, start_tok: int,
                       max_predicts: int, alpha: float) -> List[int]:
    """samd_sam_only/sam/dyn_sam.py:116-121: no anchor walk, no padding."""
    n = draft_budget(match_length, max_predicts, alpha)
    e = sam.first_end[state]
    return [int(start_tok)] + sam.text[e + 1:e + n]


def build_topk(sam: Automaton, k: int = 8) -> List[List[Tuple[int, int]]]:
    """samd_sam_only/sam/static_sam.py:137-146: per state, out-edges stably sorted by the
    target's occurrence count, descending; ties keep dict insertion order."""
    out = []
    for s in range(sam.n_states):
        edges = list(sam.trans[s].items())
        edges.sort(key=lambda e: -sam.occ[e[1]])      # list.sort is stable
        out.append(edges[:k])
    return out


class _Heap:
    """Binary min-heap on `key` with CPython heapq's exact sift order (Lib/heapq.py:
    heappush = append + _siftdown; heappop = move last to root, _siftup to a leaf
    preferring the right child unless left < right strictly, then _siftdown)."""

    def __init__(self):
        self.a: List[tuple] = []

    def __len__(self):
        return len(self.a)

    def _down(self, start: int, pos: int):
        a = self.a
        item = a[pos]
        while pos > start:
            parent = (pos - 1) >> 1
            if item[0] < a[parent][0]:
                a[pos] = a[parent]
                pos = parent
                continue
            break
        a[pos] = item

    def push(self, item: tuple):
        self.a.append(item)
        self._down(0, len(self.a) - 1)

    def pop(self) -> tuple:
        a = self.a
        last = a.pop()
        if not a:
            return last
        top = a[0]
        a[0] = last
        end = len(a)
        pos = 0
        item = a[0]
        child = 1
        while child < end:
            right = child + 1
            if right < end and not (a[child][0] < a[right][0]):
                child = right
            a[pos] = a[child]
            pos = child
            child = 2 * pos + 1
        a[pos] = item
        self._down(0, pos)
        return top


def static_tree_sam_only(sam: Automaton, topk: List[List[Tuple[int, int]]], state: int, match_length: int,
                         start_tok: int, max_predicts: int, alpha: float, K: int):
    """samd_sam_only/sam/static_sam.py:182-215: best-first tree over occurrence-count
    ratios.  `match_length` is already biased by the caller.  Returns (tokens, parents)."""
    n = draft_budget(match_length, max_predicts, alpha)
    heap = _Heap()
    tokens: List[int] = []
    parents: List[int] = []
    per_depth: Dict[int, int] = {}
    heap.push((-1.0, int(start_tok), state, -1, 0))
    while len(tokens) != n and len(heap) != 0:
        prob, tok, st, par, depth = heap.pop()
        if per_depth.get(depth, 0) + 1 > K:
            per_depth.setdefault(depth, 0)
            continue
        per_depth[depth] = per_depth.get(depth, 0) + 1
        me = len(tokens)
        tokens.append(tok)
        parents.append(par)
        if len(tokens) == n:
            break
        total = sam.occ[st]
        for ntok, nst in topk[st][:K]:
            ratio = sam.occ[nst] / total               # division first, then multiply
            heap.push((prob * ratio, ntok, nst, me, depth + 1))
    return tokens, parents


def tree_buffers(parents: Sequence[int]):
    """samd_sam_only/sam/static_sam.py:148-180: (mask [n,n] bool, depth [n], retrieve
    [leaves, maxdepth] with -1 padding; leaves in ascending node order)."""
    n = len(parents)
    leaf = [True] * n
    depth = [0] * n
    for i in range(1, n):
        leaf[parents[i]] = False
        depth[i] = depth[parents[i]] + 1
    mask = np.zeros((n, n), dtype=bool)
    for i in range(n):
        j = i
        while j != -1:
            mask[i, j] = True
            j = parents[j]
    paths = []
    for i in range(n):
        if not leaf[i]:
            continue
        p = [i]
        while p[-1] != 0:
            p.append(parents[p[-1]])
        paths.append(p[::-1])
    width = max(len(p) for p in paths)
    retrieve = np.array([p + [-1] * (width - len(p)) for p in paths], dtype=np.int64)
    return mask, np.array(depth, dtype=np.int64), retrieve


def select_sam_only(dyn: Automaton, static: Automaton, topk, start_tok: int, max_predicts: int, alpha: float,
                    K: int, len_bias: int):
    """samd_sam_only/draft.py:49-59: dyn sequence when match_dyn >= match_static - len_bias,
    else static tree."""
    si, sl = dyn.peek(start_tok)
    ti, tl = static.peek(start_tok)
    tl_b = tl - len_bias
    info = dict(index_dyn=si, match_dyn=sl, index_static=ti, match_static=tl)
    if sl >= tl_b:
        return "sequence", dyn_draft_sam_only(dyn, si, sl, start_tok, max_predicts, alpha), None, info
    toks, par = static_tree_sam_only(static, topk, ti, tl_b, start_tok, max_predicts, alpha, K)
    return "tree", toks, par, info


# --------------------------------------------------------------------------------------
# Static corpus construction
# --------------------------------------------------------------------------------------
def build_static(docs: Iterable[Sequence[int]], eos: int, count_occurrences: bool = False) -> Automaton:
    """samd/sam/static_sam.py:32-46 (+ samd_sam_only/sam/static_sam.py:31-39,126-130):
    every document is appended, followed by EOS unless it already ends with it.  The
    cursor also moves during the build (static_sam.py:96-100) but `reset()` clears it
    before use."""
    sam = Automaton(count_occurrences=count_occurrences, keep_text=not count_occurrences)
    for d in docs:
        d = [int(t) for t in d]
        sam.extend(d)
        if d[-1] != eos:
            sam.extend([eos])
    sam.reset_cursor()
    return sam


# --------------------------------------------------------------------------------------
# Greedy verification + KV compaction
# --------------------------------------------------------------------------------------
def row_argmax(logits) -> np.ndarray:
    """torch.argmax contract used at samd/utils.py:86,131: lowest index among equal
    maxima, NaN counts as the maximum (first NaN wins), +0.0 == -0.0.  `logits` is
    [..., V] (numpy float, or a torch tensor of any float dtype - converted exactly)."""
    try:
        import torch
        if isinstance(logits, torch.Tensor):
            logits = logits.detach().to("cpu", torch.float32).numpy()
    except ImportError:  # pragma: no cover
        pass
    x = np.asarray(logits, dtype=np.float32)
    nan = np.isnan(x)
    has_nan = nan.any(axis=-1)
    first_nan = nan.argmax(axis=-1)
    plain = np.where(nan, -np.inf, x).argmax(axis=-1)
    return np.where(has_nan, first_nan, plain).astype(np.int64)


def row_topk(logits, k: int = 8) -> np.ndarray:
    """torch.topk(k).indices per row as used at samd/tree_model/token_recycle/token_recycle.py:36-38, under the
    deterministic refinement (value descending, index ascending): torch leaves the order - and the choice - of
    equal values unspecified, so equality with the reference is defined on rows whose k+1 largest values are
    distinct, and on the multiset of selected VALUES otherwise.  NaN counts as the largest value (torch.topk's
    rule, same as argmax), +0.0 == -0.0.  Column 0 is row_argmax()."""
    try:
        import torch
        if isinstance(logits, torch.Tensor):
            logits = logits.detach().to("cpu", torch.float32).numpy()
    except ImportError:  # pragma: no cover
        pass
    x = np.asarray(logits, dtype=np.float32)
    key = x.astype(np.float64)
    key = np.where(np.isposinf(key), np.finfo(np.float64).max, key)
    key = np.where(np.isnan(x), np.inf, key)                    # NaN above +inf
    order = np.argsort(-key, axis=-1, kind="stable")            # stable: equal values keep index order
    return order[..., :k].astype(np.int64)


def recycle_update(cache: dict, tree_tokens, topk_rows) -> None:
    """TokenRecycle.update (token_recycle.py:39-47): cache[token] = that row's top-k, rows in order - a token
    that occurs twice keeps the LAST row's list."""
    for tok, best in zip(np.asarray(tree_tokens).tolist(), np.asarray(topk_rows).tolist()):
        cache[int(tok)] = [int(b) for b in best]


def recycle_gen_draft(cache: dict, tree: List[List[int]], start_token: int) -> List[int]:
    """TokenRecycle.gen_draft (token_recycle.py:49-59): fill the static tree top-down; the children of a node
    whose token has no entry keep their current value (0)."""
    tokens = [int(start_token)] + [0] * (len(tree) - 1)
    for node, childs in enumerate(tree):
        best = cache.get(tokens[node])
        if best is None:
            continue
        for j, child in enumerate(childs):
            tokens[child] = best[j]
    return tokens


def verify_greedy(node_argmax: np.ndarray, tree_tokens: np.ndarray, retrieve: np.ndarray):
    """samd/samd_model.py:159-168 (gather) + samd/utils.py:127-141 (greedy posterior) +
    samd/samd_model.py:195-199 (accepted tokens / indices) for ONE request.

    node_argmax [T] : argmax of every tree row;  tree_tokens [T];  retrieve [P, D] with
    -1 padding.  A -1 entry selects the appended 0 token (samd/utils.py:95-96) and, in
    the logits gather, wraps to the last row (samd/samd_model.py:164).
    Returns dict(best, accept_len (= accepted count + 1), next_token, tokens, indices).
    """
    T = len(tree_tokens)
    P, D = retrieve.shape
    ext = np.concatenate([np.asarray(tree_tokens, dtype=np.int64), [0]])
    cand = ext[retrieve]                                  # -1 -> appended 0
    rows = np.where(retrieve < 0, T - 1, retrieve)        # -1 -> last row
    hit = (cand[:, 1:] == node_argmax[rows[:, :-1]]).astype(np.int64)
    acc_p = np.cumprod(hit, axis=1).sum(axis=1) if D > 1 else np.zeros(P, dtype=np.int64)
    acc = int(acc_p.max()) if P else 0
    best = 0 if acc == 0 else int(acc_p.argmax())          # first maximal path
    return dict(best=best, accept_len=acc + 1, next_token=int(node_argmax[rows[best, acc]]),
                tokens=cand[best, :acc + 1].astype(np.int64), indices=retrieve[best, :acc + 1].astype(np.int64),
                path_accept=acc_p)


def verify_sequence(node_argmax: np.ndarray, seq_tokens: np.ndarray):
    """Sequence drafts: candidates [1, n] = the draft, logits [1, n, V]
    (samd/samd_model.py:160-162): one path, identity retrieve, no KV move."""
    n = len(seq_tokens)
    ident = np.arange(n, dtype=np.int64)[None, :]
    r = verify_greedy(node_argmax, seq_tokens, ident)
    r["indices"] = None
    return r


def kv_compact(kv: List[np.ndarray], start: int, indices: Optional[np.ndarray], accept_len: int) -> int:
    """samd/cache.py:118-133 for one request: every tensor [H, max_len, Dh]; rows
    start+indices[j] -> start+j (gather to a temporary, then copy).  Returns the new
    cache length.  Sequence drafts (indices None) only bump the length."""
    if indices is not None:
        src = start + np.asarray(indices, dtype=np.int64)
        for t in kv:
            tmp = t[..., src, :].copy()
            t[..., start:start + accept_len, :] = tmp
    return start + accept_len


# --------------------------------------------------------------------------------------
# Fake-LM decode loop (restates the glue of samd/samd_model.py:131-274 without the LLM)
# --------------------------------------------------------------------------------------
def truth_next(tok: int, vocab: int) -> int:
    """Deterministic fake language model: the 'true' next token after `tok`."""
    return (tok * 7 + 3) % vocab


def generate_sam_only(prompt: Sequence[int], truth: Sequence[int], max_new_tokens: int, max_predicts: int,
                      alpha: float, len_bias: int = 5, static: Optional[Automaton] = None, topk=None, K: int = 8,
                      eos: Optional[int] = None):
    """Greedy decode loop of samd_sam_only/samd_model.py:96-237 with the LM replaced by
    an oracle that knows the continuation `truth` (the token following absolute position
    i is truth[i]).  Only sequence drafts are verified here (static tree needs the tree
    path too; covered by verify_greedy tests).  Returns (output tokens, accept lengths)."""
    dyn = Automaton()
    dyn.extend(prompt)
    if static is not None:
        static.reset_cursor()
        static.advance(prompt)
    out: List[int] = []
    accepts: List[int] = []
    pos = len(prompt)
    start_tok = int(truth[pos])
    while len(out) < max_new_tokens and pos + max_predicts < len(truth) - 1:
        si, sl = dyn.peek(start_tok)
        draft = dyn_draft_sam_only(dyn, si, sl, start_tok, max_predicts, alpha)
        # the fake LM's argmax after consuming draft[:j+1] at absolute position pos+j is truth[pos+j+1]
        acc = 0
        while acc + 1 < len(draft) and draft[acc + 1] == int(truth[pos + acc + 1]):
            acc += 1
        accepted = draft[:acc + 1]
        next_tok = int(truth[pos + acc + 1])
        dyn.extend(accepted)
        if static is not None:
            static.advance(accepted)
        if eos is not None and eos in accepted:
            accepted = accepted[:accepted.index(eos) + 1]
            out.extend(accepted)
            accepts.append(len(accepted))
            break
        out.extend(accepted)
        accepts.append(len(accepted))
        pos += len(accepted)
        start_tok = next_tok
    return out[:max_new_tokens], accepts


# --------------------------------------------------------------------------------------
# Brute-force definition (second, independent oracle for small cases)
# --------------------------------------------------------------------------------------
def brute_peek(history: Sequence[int], tok: int) -> Tuple[int, int]:
    """(L, e): L = length of the longest suffix of history+[tok] that occurs inside
    history; e = 1-based end position of its earliest occurrence; (0, 0) if none."""
    h = list(history)
    s = h + [int(tok)]
    n = len(h)
    for L in range(min(len(s), n), 0, -1):
        pat = s[len(s) - L:]
        for end in range(L, n + 1):
            if h[end - L:end] == pat:
                return L, end
    return 0, 0


# --------------------------------------------------------------------------------------
# Stochastic (typical-acceptance) verification, samd/utils.py:142-184 + the draw of the next token (:85-88)
# --------------------------------------------------------------------------------------
def philox_uniform(seed: int, counter: int) -> float:
    """The product's RNG contract (include/samd_b200.h): Philox4x32-10, counter = {lo32(c), hi32(c), 0, 0},
    key = {lo32(seed), hi32(seed)}; u = (word 0 >> 8) * 2^-24."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c = [counter & MASK, (counter >> 32) & MASK, 0, 0]
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return float(np.float32(c[0] >> 8) * np.float32(1.0 / 16777216.0))


def processed_probs(row: np.ndarray, temperature: float, top_p: float, top_k: int) -> np.ndarray:
    """softmax(logits_processor(row)) with the processor list of SamdGenerationConfig.prepare_logits_processor
    (samd/utils.py:44-58): TemperatureLogitsWarper, then TopPLogitsWarper, then TopKLogitsWarper (transformers)."""
    s = row.astype(np.float64)
    if temperature >= 1e-5 and temperature != 1.0:
        s = s / temperature
    if 1e-8 <= top_p < 1.0:
        order = np.argsort(s, kind="stable")                       # ascending
        e = np.exp(s[order] - s.max())
        cum = np.cumsum(e / e.sum())
        remove = cum <= (1.0 - top_p)
        remove[-1] = False                                          # min_tokens_to_keep = 1
        s = s.copy()
        s[order[remove]] = -np.inf
    if top_k > 0:
        k = min(top_k, s.size)
        kth = np.sort(s)[-k]
        s = np.where(s < kth, -np.inf, s)
    e = np.exp(s - s.max())
    return e / e.sum()


def verify_typical(logits: np.ndarray, tree_tokens: np.ndarray, retrieve: np.ndarray, temperature: float, top_p: float,
                   top_k: int, uniform):
    """eval_posterior's sampling branch for ONE request.  logits [T, V], tree_tokens [T], retrieve [P, D] (-1 padded);
    `uniform()` supplies the draws in order (the reference calls random.random(); the product its Philox stream).
    Returns dict(best, accept_len, next_token, tokens, indices, sample_p, margins) - `margins` lists |r - p| of every
    decision so that a comparison with float32 arithmetic can skip the razor-edge ones."""
    T, V = logits.shape
    ext = np.concatenate([np.asarray(tree_tokens, dtype=np.int64), [0]])
    cand = ext[retrieve]                                            # -1 picks the appended 0
    rows = np.where(retrieve < 0, T - 1, retrieve)                  # -1 wraps to the last row
    P, D = cand.shape
    accept, best, adjust = 1, 0, False
    gtp, margins = None, []
    for i in range(1, D):
        if i != accept:
            break
        adjust = False
        is_eq = (cand[:, :accept] == cand[best, :accept]).all(axis=1)
        fi = int(np.nonzero(is_eq)[0][0])
        gtp = processed_probs(logits[rows[fi, i - 1]], temperature, top_p, top_k)
        seen = []
        for j in range(P):
            if not is_eq[j]:
                continue
            x = int(cand[j, i])
            if x in seen or x == -1:
                continue
            seen.append(x)
            r = uniform()
            margins.append(abs(r - gtp[x]))
            if r <= gtp[x]:
                accept += 1
                best = j
                break
            gtp = gtp.copy()
            gtp[x] = 0.0
            gtp = gtp / gtp.sum()
            adjust = True
    if adjust and accept != D:
        sample_p = gtp
    else:
        row = logits[rows[best, accept - 1]].astype(np.float64)
        e = np.exp(row - row.max())
        sample_p = e / e.sum()
    u = uniform()
    cdf = np.cumsum(sample_p)
    nxt = int(min(np.searchsorted(cdf, u * cdf[-1], side="right"), V - 1))
    margins.append(float(np.min(np.abs(cdf / cdf[-1] - u))))
    return dict(best=best, accept_len=accept, next_token=nxt, tokens=cand[best, :accept].tolist(),
                indices=retrieve[best, :accept].tolist(), sample_p=sample_p, margins=margins)
