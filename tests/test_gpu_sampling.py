"""Stochastic (typical-acceptance) verification, SURVEY 8f row 4: samd_verify_sample against
  (1) the oracle's restatement of samd/utils.py:142-184 fed with the SAME Philox stream - decision for decision
      (decisions closer than 1e-4 to their threshold are not compared: the kernel's arithmetic is float32), and
  (2) the REFERENCE's own eval_posterior + torch.multinomial statistics over >= 10^5 trials (tests/golden/sampling.npz,
      produced by oracle/gen_sampling_golden.py) - acceptance-length histogram, best-path histogram and the joint
      (acceptance length, next token) table, every cell within binomial error."""
import numpy as np
import pytest
import torch

import samd_oracle as O
from helpers import load

pytestmark = pytest.mark.gpu


def _mods():
    from samd_b200 import engine, _cabi, synth
    return engine, _cabi, synth


def _i32(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).cuda()


@pytest.mark.parametrize("dt,temp,top_p,top_k", [("fp32", 1.0, 0.0, 0), ("fp32", 0.7, 0.9, 0), ("fp32", 1.3, 0.0, 5),
                                                   ("bf16", 0.8, 0.85, 8), ("fp16", 1.0, 0.5, 3)])
def test_sampling_decisions_match_oracle_given_the_same_stream(dt, temp, top_p, top_k):
    E, K, synth = _mods()
    tree = synth.token_recycle_tree()
    ri = synth.tree_retrieve_indices(tree)                       # [30, 6], -1 padded
    B, T, V = 48, len(tree), 997
    rng = np.random.default_rng(11)
    logits = (rng.standard_normal((B, T, V)) * 2.0).astype(np.float32)
    tokens = rng.integers(0, V, size=(B, T)).astype(np.int32)
    parent = {}
    for n, ch in enumerate(tree):
        for c in ch:
            parent[c] = n
    for b in range(B):                                           # proposals that are often accepted, and sibling duplicates
        for c, p in parent.items():
            if rng.random() < 0.8:
                logits[b, p, tokens[b, c]] += 4.0
        tokens[b, 2] = tokens[b, 1]
    tdt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    lg = torch.as_tensor(logits).to(tdt).cuda()
    lg_np = lg.float().cpu().numpy()                             # what the kernel really reads
    seeds = torch.as_tensor(rng.integers(1, 2 ** 62, size=B)).cuda()
    off0 = rng.integers(0, 2 ** 40, size=B)
    offs = torch.as_tensor(off0.copy()).cuda()
    ver = E.Verifier(B, T)
    out = ver.verify_sample(lg, _i32(tokens), _i32(ri), temp, top_p, top_k, seeds, offs, want_sample_p=True)
    torch.cuda.synchronize()
    compared = 0
    for b in range(B):
        k = [0]

        def draw():
            u = O.philox_uniform(int(seeds[b]), int(off0[b]) + k[0])
            k[0] += 1
            return u

        want = O.verify_typical(lg_np[b], tokens[b], ri.astype(np.int64), temp, top_p, top_k, draw)
        if min(want["margins"]) < 1e-4:
            continue
        compared += 1
        assert out["best"][b].item() == want["best"] and out["accept_len"][b].item() == want["accept_len"], b
        n = want["accept_len"]
        assert out["tokens"][b, :n].tolist() == want["tokens"] and out["indices"][b, :n].tolist() == want["indices"], b
        assert (out["tokens"][b, n:] == -1).all()
        assert int(offs[b]) - int(off0[b]) == k[0], b            # the stream advanced by exactly the draws used
        assert np.allclose(out["sample_p"][b].cpu().numpy(), want["sample_p"], rtol=2e-3, atol=1e-6), b
        assert out["next_token"][b].item() == want["next_token"], b
    assert compared >= B * 3 // 4


def test_sampling_sequence_candidates():
    """A sequence draft is one identity path (retrieve = None)."""
    E, K, synth = _mods()
    B, T, V = 16, 12, 300
    rng = np.random.default_rng(5)
    logits = (rng.standard_normal((B, T, V)) * 2.0).astype(np.float32)
    tokens = rng.integers(0, V, size=(B, T)).astype(np.int32)
    for b in range(B):
        for t in range(T - 1):
            logits[b, t, tokens[b, t + 1]] += 5.0
    seeds = torch.arange(1, B + 1, dtype=torch.int64).cuda()
    offs = torch.zeros(B, dtype=torch.int64).cuda()
    out = E.Verifier(B, T).verify_sample(torch.as_tensor(logits).cuda(), _i32(tokens), None, 0.9, 0.0, 0, seeds, offs)
    torch.cuda.synchronize()
    ident = np.arange(T, dtype=np.int64)[None, :]
    for b in range(B):
        k = [0]

        def draw():
            u = O.philox_uniform(b + 1, k[0])
            k[0] += 1
            return u

        want = O.verify_typical(logits[b], tokens[b], ident, 0.9, 0.0, 0, draw)
        if min(want["margins"]) < 1e-4:
            continue
        assert out["accept_len"][b].item() == want["accept_len"] and out["next_token"][b].item() == want["next_token"]


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_sampling_statistics_match_reference(ci):
    E, K, synth = _mods()
    z = load("sampling.npz")
    temp, top_p, top_k = z["configs"][ci]
    trials = int(z["trials"])
    logits, tokens, ri = z["logits"], z["tree_tokens"], z["retrieve"]
    T, V = logits.shape
    P, D = ri.shape
    B = 20000
    lg = torch.as_tensor(logits).cuda()[None].expand(B, T, V)    # every trial sees the same problem (batch stride 0)
    tk = _i32(np.tile(tokens[None], (B, 1)))
    ver = E.Verifier(B, T)
    h_acc, h_best, h_joint = np.zeros(D + 1, np.int64), np.zeros(P, np.int64), np.zeros((D + 1, V), np.int64)
    done = 0
    while done < trials:
        n = min(B, trials - done)
        seeds = (torch.arange(B, dtype=torch.int64) + 1 + done + 7919 * (ci + 1) * 1000003).cuda()
        offs = torch.zeros(B, dtype=torch.int64).cuda()
        out = ver.verify_sample(lg, tk, _i32(ri), float(temp), float(top_p), int(top_k), seeds, offs)
        acc = out["accept_len"][:n].cpu().numpy()
        h_acc += np.bincount(acc, minlength=D + 1)
        h_best += np.bincount(out["best"][:n].cpu().numpy(), minlength=P)
        np.add.at(h_joint, (acc, out["next_token"][:n].cpu().numpy()), 1)
        done += n

    def close(got, ref, what):
        # two samples of `trials` draws each: the difference of two binomial counts has variance <= got + ref
        tol = 5.0 * np.sqrt(got + ref + 1.0) + 3.0
        bad = np.abs(got - ref) > tol
        assert not bad.any(), (what, np.argwhere(bad)[:5].tolist(), got[bad][:5].tolist(), ref[bad][:5].tolist())

    close(h_acc, z[f"c{ci}/accept_hist"], "accept length")
    close(h_best, z[f"c{ci}/best_hist"], "best path")
    close(h_joint, z[f"c{ci}/joint"], "accept length x next token")
