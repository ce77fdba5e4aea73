"""Synthetic token-stream generators for the SAM hot path.

These are workload generators only (no algorithmic content): tests, the oracle
fixtures script and bench.py all draw their inputs from here so that the CUDA
path, the oracle and the reference see the same seeded streams.

Token ids 0,1,2 are reserved (pad / BOS / EOS; EOS = 2 as in Llama/Vicuna).
"""
from __future__ import annotations

import numpy as np

EOS = 2
FIRST_TOKEN = 3


def _fresh_tokens(rng: np.random.Generator, k: int, vocab: int, zipf_a: float, uniform: bool) -> np.ndarray:
    if uniform or vocab <= 8:
        return rng.integers(min(FIRST_TOKEN, vocab - 1) if vocab > FIRST_TOKEN else 0, vocab, size=k, dtype=np.int64)
    z = rng.zipf(zipf_a, size=k)
    return np.clip(z + (FIRST_TOKEN - 1), FIRST_TOKEN, vocab - 1).astype(np.int64)


def copy_mix(n: int, vocab: int, seed: int, p_copy: float = 0.5, span=(4, 32), zipf_a: float = 1.2,
             source: np.ndarray | None = None, p_source: float = 0.0, uniform_fresh: bool = False) -> np.ndarray:
    """Stream of `n` tokens: with probability `p_copy` append a span copied from an
    earlier offset of the same stream (or, with probability `p_source`, from
    `source` - e.g. a static corpus), else one fresh token (Zipf(zipf_a) clipped to
    [3, vocab), or uniform when `uniform_fresh` - the clone-light regime)."""
    rng = np.random.default_rng(seed)
    out = np.empty(n + span[1] + 1, dtype=np.int64)
    pos = 0
    # warm start: a few fresh tokens so that copies have something to copy from
    k0 = min(n, 8)
    out[:k0] = _fresh_tokens(rng, k0, vocab, zipf_a, uniform_fresh)
    pos = k0
    while pos < n:
        if rng.random() < p_copy:
            length = int(rng.integers(span[0], span[1] + 1))
            if source is not None and len(source) > length and rng.random() < p_source:
                off = int(rng.integers(0, len(source) - length))
                out[pos:pos + length] = source[off:off + length]
            else:
                off = int(rng.integers(0, pos))
                length = min(length, pos - off)  # never read past what exists
                out[pos:pos + length] = out[off:off + length]
            pos += length
        else:
            out[pos] = _fresh_tokens(rng, 1, vocab, zipf_a, uniform_fresh)[0]
            pos += 1
    return out[:n].copy()


def small_alphabet(n: int, vocab: int, seed: int) -> np.ndarray:
    """Adversarial clone-heavy stream: uniform over a tiny alphabet [0, vocab)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, vocab, size=n, dtype=np.int64)


def make_corpus(n_tokens: int, vocab: int, seed: int, doc_len=(64, 512), pool_tokens: int | None = None,
                eos: int = EOS, singletons: bool = True):
    """Static-SAM corpus per the reference convention (tools/gen_sam_alpaca.py:39-44):
    documents (EOS appended by the builder) followed by one single-token document
    per vocab id.  Documents are cut from a shared copy_mix "phrase pool" so that
    cross-document repeats exist.  Returns a list of int64 arrays."""
    rng = np.random.default_rng(seed)
    if pool_tokens is None:
        pool_tokens = max(1024, n_tokens // 12)
    pool = copy_mix(pool_tokens, vocab, seed + 7, p_copy=0.3)
    docs = []
    total = 0
    while total < n_tokens:
        length = int(rng.integers(doc_len[0], doc_len[1] + 1))
        length = min(length, n_tokens - total) or 1
        off = int(rng.integers(0, max(1, pool_tokens - length)))
        d = pool[off:off + length].copy()
        # sprinkle noise so documents are not pure substrings of the pool
        m = rng.random(len(d)) < 0.02
        d[m] = rng.integers(FIRST_TOKEN, vocab, size=int(m.sum()))
        d[d == eos] = FIRST_TOKEN
        docs.append(d)
        total += len(d) + 1
    if singletons:
        docs.extend(np.array([i], dtype=np.int64) for i in range(vocab))
    return docs


def corpus_queries(docs, n_queries: int, length: int, vocab: int, seed: int, p_corpus: float = 0.7) -> np.ndarray:
    """[n_queries, length] token windows: p_corpus drawn from corpus documents, rest noise."""
    rng = np.random.default_rng(seed)
    big = [d for d in docs if len(d) > 1]
    flat = np.concatenate(big) if big else np.zeros(1, dtype=np.int64)
    out = rng.integers(FIRST_TOKEN, vocab, size=(n_queries, length), dtype=np.int64)
    for q in range(n_queries):
        pos = 0
        while pos < length:
            span = int(rng.integers(4, 33))
            span = min(span, length - pos)
            if rng.random() < p_corpus and len(flat) > span:
                off = int(rng.integers(0, len(flat) - span))
                out[q, pos:pos + span] = flat[off:off + span]
            pos += span
    return out


def token_recycle_tree():
    """The 61-node static draft tree of samd/config/token_recycle.json (root + 60),
    as children lists in BFS numbering.  Data, not code: config 4's tree shape."""
    adj = {0: [1, 2, 3, 4, 5, 6, 7], 1: [8, 9, 10, 11, 12, 13], 2: [14, 15, 16, 17, 18], 3: [19, 20, 21],
           4: [22, 23], 5: [24, 25], 6: [26], 7: [27], 8: [28, 29, 30], 9: [31, 32], 10: [33, 34], 11: [35],
           12: [36], 14: [37, 38, 39], 15: [40, 41], 16: [42], 17: [43], 19: [44], 20: [45], 22: [46],
           24: [47], 26: [48], 28: [49, 50], 29: [51], 31: [52], 37: [53, 54], 40: [55], 44: [56],
           49: [57, 58], 51: [59], 53: [60]}
    return [adj.get(i, []) for i in range(61)]


def tree_retrieve_indices(tree, reverse_leaves: bool = True) -> np.ndarray:
    """Root-to-leaf paths of a children-list tree, -1 padded to the max depth.
    Token-Recycle lists leaves in reversed node order
    (samd/tree_model/token_recycle/utils.py:77-90)."""
    parent = {0: -1}
    for node, childs in enumerate(tree):
        for c in childs:
            parent[c] = node
    paths = []
    for node, childs in enumerate(tree):
        if childs:
            continue
        p = [node]
        while p[-1] != 0:
            p.append(parent[p[-1]])
        paths.append(p[::-1])
    if reverse_leaves:
        paths = paths[::-1]
    depth = max(len(p) for p in paths)
    return np.array([p + [-1] * (depth - len(p)) for p in paths], dtype=np.int32)


def planted_logits(batch: int, n_nodes: int, vocab: int, tree_tokens: np.ndarray, retrieve: np.ndarray,
                   seed: int, dtype="bfloat16", plant: float = 20.0, device="cpu"):
    """Random logits [batch, n_nodes, vocab] with a planted accepted prefix per request
    (SURVEY.md §8d C4): for request b pick a path p and depth a in [0, D-1]; make
    tree_tokens[b, ri[p, j+1]] the argmax of row ri[p, j] for j < a."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    logits = torch.randn(batch, n_nodes, vocab, generator=g, dtype=torch.float32)
    rng = np.random.default_rng(seed + 1)
    n_paths, depth = retrieve.shape
    planted = []
    for b in range(batch):
        p = int(rng.integers(0, n_paths))
        real = int((retrieve[p] >= 0).sum())
        a = int(rng.integers(0, real))
        for j in range(a):
            logits[b, retrieve[p, j], int(tree_tokens[b, retrieve[p, j + 1]])] = plant
        planted.append((p, a))
    logits = logits.to(getattr(torch, dtype))
    if device != "cpu":
        logits = logits.to(device)
    return logits, planted
