#!/usr/bin/env python
"""Statistics of the REFERENCE's stochastic verification (samd/utils.py:142-184 eval_posterior with greedy=False, then
torch.multinomial(sample_p, 1) as gen_candidates does, :85-88) on a small fixed problem: run the reference's own function
N times per configuration and store the histograms -> tests/golden/sampling.npz.  Build container only (imports
/root/reference through oracle/ref_loader.py).  The product's Philox stream cannot equal Python's random.random(), so
parity with the reference is statistical: tests/test_gpu_sampling.py compares the CUDA kernel's histograms over the same
number of trials with these, within binomial error.

    python oracle/gen_sampling_golden.py [trials]
"""
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

DET_TRIALS = 300
CONFIGS = [(1.0, 0.0, 0), (0.7, 0.9, 0), (1.3, 0.0, 5), (0.8, 0.85, 8)]      # (temperature, top_p, top_k)


def problem():
    """13-node tree, V = 24.  Children lists in BFS numbering; tokens chosen so that siblings repeat a token once (the
    candidates_set skip) and a padded path exists (-1 -> token 0 / last row)."""
    tree = [[1, 2, 3], [4, 5], [6], [], [7, 8], [], [9], [10], [], [11], [], [12], []]
    T, V = len(tree), 24
    rng = np.random.default_rng(77)
    logits = (rng.standard_normal((T, V)) * 1.5).astype(np.float32)
    tokens = rng.integers(1, V, size=T).astype(np.int64)
    tokens[2] = tokens[1]                                     # two children of the root propose the same token
    parent = {0: -1}
    for n, ch in enumerate(tree):
        for c in ch:
            parent[c] = n
    for n, ch in enumerate(tree):                             # make most proposals likely: boost the child's token in the parent row
        for c in ch:
            logits[n, tokens[c]] += 2.5
    paths = []
    for n, ch in enumerate(tree):
        if not ch:
            p = [n]
            while p[-1] != 0:
                p.append(parent[p[-1]])
            paths.append(p[::-1])
    paths = paths[::-1]
    D = max(len(p) for p in paths)
    ri = np.array([p + [-1] * (D - len(p)) for p in paths], dtype=np.int64)
    return logits, tokens, ri


def main(trials):
    ns = ref_loader.load()
    logits, tokens, ri = problem()
    T, V = logits.shape
    P, D = ri.shape
    lt = torch.tensor(logits)
    ext = torch.tensor(np.concatenate([tokens, [0]]))
    cand = ext[torch.tensor(ri)]
    gathered = lt[torch.tensor(ri)]                           # [P, D, V]; -1 wraps to the last row
    out = dict(logits=logits, tree_tokens=tokens.astype(np.int32), retrieve=ri.astype(np.int32), trials=np.array(trials),
               configs=np.array(CONFIGS, dtype=np.float64))
    for ci, (temp, top_p, top_k) in enumerate(CONFIGS):
        with contextlib.redirect_stdout(io.StringIO()):
            cfg = ns.samd_utils.SamdGenerationConfig(greedy=False, temperature=temp, top_p=top_p, top_k=int(top_k))
        random.seed(1000 + ci)
        torch.manual_seed(2000 + ci)
        h_acc = np.zeros(D + 1, dtype=np.int64)
        h_best = np.zeros(P, dtype=np.int64)
        h_joint = np.zeros((D + 1, V), dtype=np.int64)
        for _ in range(trials):
            best, acc, sample_p = ns.samd_utils.eval_posterior(gathered, cand, cfg)
            nxt = int(torch.multinomial(sample_p, 1).item())
            h_acc[int(acc)] += 1
            h_best[int(best)] += 1
            h_joint[int(acc), nxt] += 1
        out[f"c{ci}/accept_hist"], out[f"c{ci}/best_hist"], out[f"c{ci}/joint"] = h_acc, h_best, h_joint
        print(ci, (temp, top_p, top_k), "accept hist", h_acc.tolist())
        # deterministic section: the reference's random.random() replaced by a recorded stream (the product's Philox
        # contract, seed 5000 + ci, consecutive counters), so that the oracle's restatement can be pinned decision by
        # decision: per trial the draws consumed, best, accept length and the returned sample_p
        import samd_oracle as O
        state = {"k": 0}

        def draw():
            u = O.philox_uniform(5000 + ci, state["k"])
            state["k"] += 1
            return u

        real = ns.samd_utils.random.random
        ns.samd_utils.random.random = draw
        det = []
        sps = []
        try:
            for _ in range(DET_TRIALS):
                k0 = state["k"]
                best, acc, sample_p = ns.samd_utils.eval_posterior(gathered, cand, cfg)
                det.append((k0, state["k"] - k0, int(best), int(acc)))
                sps.append(sample_p.view(-1).numpy().astype(np.float32))
        finally:
            ns.samd_utils.random.random = real
        out[f"c{ci}/det"] = np.array(det, dtype=np.int64)
        out[f"c{ci}/det_sample_p"] = np.stack(sps)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "sampling.npz"), **out)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000)
