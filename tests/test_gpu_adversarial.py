"""GPU parity on structured worst-case streams (tests/adversarial.py) and hypothesis-generated small alphabets:
DynSamBatch through the C ABI against the oracle, STATE FOR STATE - numbering, links, lengths, min_endpos, every
state's edges in insertion order, cursor, lookups and drafts after every chunk - for both variants of the step
kernel (1 = one thread per request, 0 = the warp-cooperative probe with its lane-parallel chain insert).
Reference: /root/reference/samd/sam/dyn_sam.py:41-113."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import samd_oracle as O
import adversarial as A

pytestmark = pytest.mark.gpu
BIG = 1 << 20


def _mods():
    from samd_b200 import engine, _cabi
    return engine, _cabi


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).cuda()


def run_batch(cases, variant, probes_of, n_predicts=16, check_every=1):
    """cases: list of (stream, chunk sizes).  All requests advance in lock step (ragged counts)."""
    E, K = _mods()
    K.lib().samd_step_set_variant(variant)
    try:
        B = len(cases)
        cap = max(len(s) for s, _ in cases) + 8
        dyn = E.DynSamBatch(B, cap)
        eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=n_predicts, len_bias=0, len_threshold=-BIG)
        refs = [O.Automaton() for _ in range(B)]
        pos = [0] * B
        n_steps = max(len(c) for _, c in cases)
        for step in range(n_steps):
            cnt = np.array([c[step] if step < len(c) else 0 for _, c in cases], dtype=np.int32)
            tok = np.zeros((B, max(1, int(cnt.max()))), dtype=np.int32)
            for r, (s, _) in enumerate(cases):
                tok[r, :cnt[r]] = s[pos[r]:pos[r] + cnt[r]]
                refs[r].extend(s[pos[r]:pos[r] + cnt[r]])
                pos[r] += cnt[r]
            eng.step(_dev(tok), _dev(cnt), None)
            if step % check_every and step != n_steps - 1:
                continue
            for j in range(max(len(probes_of(s)) for s, _ in cases)):
                start = np.array([probes_of(s)[j % len(probes_of(s))] for s, _ in cases], dtype=np.int32)
                eng.step(None, None, _dev(start))
                torch.cuda.synchronize()
                idx, ml, dr = eng.index_dyn.cpu().numpy(), eng.match_dyn.cpu().numpy(), eng.draft.cpu().numpy()
                for r in range(B):
                    want = refs[r].peek(int(start[r]))
                    assert (int(idx[r]), int(ml[r])) == want, (r, step, int(start[r]))
                    assert dr[r].tolist() == O.dyn_draft_samd(refs[r], want[0], int(start[r]), n_predicts), (r, step)
        meta = dyn.meta()
        for r in range(B):
            ex, ref = dyn.export(r), refs[r]
            assert ex["overflow"] == 0
            assert (ex["n_states"], ex["max_length"], ex["n_clones"]) == (ref.n_states, ref.n, ref.n_clones), r
            assert (ex["cur_index"], ex["cur_length"], ex["last"]) == (ref.cur, ref.cur_len, ref.tail), r
            assert ex["link"].tolist() == ref.link and ex["length"].tolist() == ref.length, r
            assert ex["min_endpos"].tolist() == ref.first_end, r
            want = [(v, t, g) for v in range(ref.n_states) for t, g in ref.trans[v].items()]
            assert [tuple(e) for e in dyn.export_edges(r).tolist()] == want, r
        return meta
    finally:
        K.lib().samd_step_set_variant(1)


@pytest.mark.parametrize("variant", [1, 0])
def test_adversarial_streams_state_for_state(variant):
    streams = A.streams()
    names, cases = [], []
    for name in sorted(streams):
        s = streams[name]
        for mode, sizes in A.chunkings(len(s), 5).items():
            if mode == "single" and len(s) > 140:
                continue
            names.append((name, mode))
            cases.append((s, sizes))
    meta = run_batch(cases, variant, lambda s: sorted(set(s))[:3] + [9999], check_every=3)
    # the long fallback chains were really walked by the kernel: a^k b visits the k a-states and the root
    longest = {n: int(meta[i, 14]) for i, n in enumerate(names)}
    for k in (31, 32, 33, 63, 64, 65, 200):
        for mode in ("steps1-8", "whole", "eights"):
            assert longest[(f"a^{k} b", mode)] == k, (k, mode, longest[(f"a^{k} b", mode)])
    assert any(32 < v <= 64 for v in longest.values()) and any(v > 64 for v in longest.values())


@pytest.mark.parametrize("variant", [1, 0])
def test_one_step_of_eight_tokens_that_all_split(variant):
    """The mailbox scout and the redirect scout get a hand-off for every token of the step."""
    s = A.all_split(rounds=10)
    head = 9                               # [30] + the block, appended first
    sizes = [head] + [1, 8] * 10           # every re-entry: the new left context, then the whole block in ONE step
    assert sum(sizes) == len(s)
    ref = O.Automaton()
    ref.extend(s[:head + 1])
    before = ref.n_clones
    ref.extend(s[head + 1:head + 9])
    assert ref.n_clones - before >= 7      # (the oracle confirms the premise: >= 7 of the 8 tokens split a state)
    run_batch([(s, sizes)] * 3, variant, lambda s: [10, 13, 17])


@pytest.mark.parametrize("variant", [1, 0])
def test_random_small_alphabets_batch(variant):
    rng = np.random.default_rng(31)
    cases = []
    for v in (2, 3, 4, 6):
        for i in range(8):
            n = int(rng.integers(200, 1200))
            s = rng.integers(3, 3 + v, size=n).tolist()
            cases.append((s, A.chunkings(n, 100 * v + i)["steps1-8"]))
    run_batch(cases, variant, lambda s: [3, 4, 5], check_every=5)


@pytest.mark.parametrize("variant", [1, 0])
def test_hypothesis_small_alphabets(variant):
    """hypothesis over V in {2,3,4}, n <= 300, chunked 1-8: 32 generated streams per batch launch."""
    pool = []

    @settings(max_examples=96, deadline=None, database=None, derandomize=True)
    @given(st.integers(2, 4).flatmap(lambda v: st.tuples(st.lists(st.integers(3, 2 + v), min_size=1, max_size=300),
                                                         st.integers(0, 2 ** 31 - 1))))
    def collect(case):
        pool.append(case)

    collect()
    for lo in range(0, len(pool), 32):
        cases = [(s, A.chunkings(len(s), seed)["steps1-8"]) for s, seed in pool[lo:lo + 32]]
        run_batch(cases, variant, lambda s: [3, 4], n_predicts=7, check_every=4)


def test_negative_token_and_full_arena_are_flagged():
    """A token of -1 would alias the free-slot marker of the layout: the kernel refuses it and raises the request's
    flag (bit 1); a full arena raises bit 0 and the cursor keeps following the text."""
    E, K = _mods()
    dyn = E.DynSamBatch(2, 16)
    eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=4, len_bias=0, len_threshold=-BIG)
    tok = np.array([[3, 4, -1, 5], [3, 4, 3, 4]], dtype=np.int32)
    eng.step(_dev(tok), None, None)
    m = dyn.meta()
    assert m[0, 6] == 2 and m[0, 2] == 2            # stopped in front of the bad token
    assert m[1, 6] == 0 and m[1, 2] == 4
    long = np.tile(np.array([[3, 4, 5, 6]], dtype=np.int32), (2, 6))[:, :20]
    eng.step(_dev(long), None, None)
    m = dyn.meta()
    assert m[1, 6] & 1 and m[1, 2] == 16 and dyn.stats()["overflowed"] >= 1


def test_lean_build_state_for_state():
    """The lean build of the step kernel (two warps, fewer registers: what batches of more than one wave run) forced on a
    small batch: ragged counts (0..8 per request and step), state for state against the oracle."""
    E, K = _mods()
    K.lib().samd_step_set_lean(1)
    try:
        rng = np.random.default_rng(77)
        cases = []
        for i in range(70):
            n = int(rng.integers(40, 400))
            s = rng.integers(3, 3 + int(rng.integers(2, 7)), size=n).tolist()
            sizes, left = [], n
            while left > 0:
                c = min(left, int(rng.integers(0, 9)))
                sizes.append(c)
                left -= c
            cases.append((s, sizes))
        run_batch(cases, 1, lambda s: [3, 4, 5], check_every=6)
        streams = A.streams()
        cases = [(streams[name], A.chunkings(len(streams[name]), 5)["steps1-8"]) for name in sorted(streams)]
        run_batch(cases, 1, lambda s: sorted(set(s))[:3] + [9999], check_every=3)
    finally:
        K.lib().samd_step_set_lean(-1)
