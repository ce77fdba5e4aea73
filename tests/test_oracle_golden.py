"""Pin the CPU oracle (oracle/samd_oracle.py) to outputs of the reference's own classes
(tests/golden/*.npz, produced by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import samd_oracle as O
from helpers import load, unragged, docs_of


@pytest.mark.parametrize("name", ["v2", "v3", "v4", "v6", "mix1k", "mix4k", "light2k"])
def test_dyn_sam_matches_reference(name):
    z = load("dyn_sam.npz")
    stream, cuts = z[f"{name}/stream"], z[f"{name}/cuts"]
    sam = O.Automaton()
    so = unragged(z[f"{name}/draft_so_flat"], z[f"{name}/draft_so_offs"])
    lo = 0
    for k, hi in enumerate(cuts):
        sam.extend(stream[lo:hi])
        lo = hi
        tok = int(stream[hi])
        i, l = sam.peek(tok)
        assert (i, l) == (z[f"{name}/index"][k], z[f"{name}/match"][k])
        assert O.dyn_draft_samd(sam, i, tok, 16) == z[f"{name}/draft16"][k].tolist()
        assert O.dyn_draft_samd(sam, i, tok, 40) == z[f"{name}/draft40"][k].tolist()
        assert O.dyn_draft_sam_only(sam, i, l, tok, 40, 4.0) == so[k]
        if len(stream) <= 256:
            L, e = O.brute_peek(stream[:hi], tok)
            assert L == l and (l == 0 or sam.first_end[i] == e)
    assert sam.link == z[f"{name}/link"].tolist()
    assert sam.length == z[f"{name}/length"].tolist()
    assert sam.first_end == z[f"{name}/min_endpos"].tolist()
    assert [sam.cur, sam.cur_len] == z[f"{name}/cursor"].tolist()


@pytest.mark.parametrize("name", ["tiny", "small", "mid"])
def test_static_sam_matches_reference(name):
    z = load("static_sam.npz")
    docs = docs_of(z, name)
    eos = int(z[f"{name}/eos"])
    a = O.build_static(docs, eos)
    b = O.build_static(docs, eos, count_occurrences=True)
    assert a.link == z[f"{name}/link"].tolist()
    assert a.length == z[f"{name}/length"].tolist()
    assert a.first_end == z[f"{name}/min_endpos"].tolist()
    assert b.occ == z[f"{name}/cnt_endpos"].tolist()
    assert a.n_edges == int(z[f"{name}/n_edges"])
    topk = O.build_topk(b, 8)
    tk_tok, tk_idx = z[f"{name}/topk_tok"], z[f"{name}/topk_idx"]
    for s, lst in enumerate(topk):
        assert [t for t, _ in lst] == [t for t in tk_tok[s] if t >= 0]
        assert [i for _, i in lst] == [i for i in tk_idx[s] if i >= 0]
    q = z[f"{name}/queries"]
    steps = z[f"{name}/steps"]
    trees = unragged(z[f"{name}/tree_tok_flat"], z[f"{name}/tree_offs"])
    depth_flat = z[f"{name}/tree_depth_flat"]
    rets = unragged(z[f"{name}/tree_ret_flat"], z[f"{name}/tree_ret_offs"])
    prev_q, pos = -1, 0
    for k, (qi, p) in enumerate(steps):
        if qi != prev_q:
            a.reset_cursor()
            b.reset_cursor()
            prev_q, pos = qi, 0
        a.advance(q[qi, pos:p])
        b.advance(q[qi, pos:p])
        pos = p
        tok = int(q[qi, p])
        i, l = a.peek(tok)
        assert (i, l) == b.peek(tok) == (z[f"{name}/index"][k], z[f"{name}/match"][k])
        assert O.static_draft_samd(a, i, tok, 16) == z[f"{name}/draft16"][k].tolist()
        toks, par = O.static_tree_sam_only(b, topk, i, max(l - 2, 0), tok, 40, 4.0, 8)
        assert toks == trees[k]
        _, depth, ret = O.tree_buffers(par)
        off = z[f"{name}/tree_offs"][k]
        assert depth.tolist() == depth_flat[off:off + len(toks)].tolist()
        assert ret.shape == tuple(z[f"{name}/tree_ret_shape"][k])
        assert ret.reshape(-1).tolist() == rets[k]


def test_draft_selection_matches_reference():
    z = load("draft_select.npz")
    docs = docs_of(z)
    sa = O.build_static(docs, 2)
    sb = O.build_static(docs, 2, count_occurrences=True)
    topk = O.build_topk(sb, 8)
    cuts_all = unragged(z["cuts_flat"], z["cuts_offs"])
    so_tok = unragged(z["so_tok_flat"], z["so_tok_offs"])
    so_ret = unragged(z["so_ret_flat"], z["so_ret_offs"])
    k = 0
    for r, stream in enumerate(z["streams"]):
        dyn = O.Automaton()
        sa.reset_cursor()
        sb.reset_cursor()
        lo = 0
        for hi in cuts_all[r]:
            chunk = stream[lo:hi]
            dyn.extend(chunk)
            sa.advance(chunk)
            sb.advance(chunk)
            lo = hi
            tok = int(stream[hi])
            typ, seq, info = O.select_samd(dyn, sa, tok, 16, 5, 5)
            src = {"dyn": 0, "static": 1, "tree": 2}[info["source"]]
            assert src == z["samd_source"][k]
            assert (0 if typ == "sequence" else 1) == z["samd_type"][k]
            if typ == "sequence":
                assert seq == z["samd_seq"][k].tolist()
            typ2, toks, par, _ = O.select_sam_only(dyn, sb, topk, tok, 40, 4.0, 8, 5)
            assert (0 if typ2 == "sequence" else 1) == z["so_type"][k]
            assert toks == so_tok[k]
            if typ2 == "tree":
                assert O.tree_buffers(par)[2].reshape(-1).tolist() == so_ret[k]
            k += 1
    assert k == len(z["samd_type"])


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_verify_matches_reference(dt):
    z = load("verify.npz")
    bits = torch.from_numpy(z["bf16/logits_bits"]).view(torch.bfloat16)
    lg = bits if dt == "bf16" else bits.float().to(torch.float16)
    am = O.row_argmax(lg)
    assert np.array_equal(am, z[f"{dt}/node_argmax"])
    ri = z["retrieve"]
    for b in range(lg.shape[0]):
        r = O.verify_greedy(am[b], z["tree_tokens"][b], ri)
        assert r["best"] == z[f"{dt}/best"][b]
        assert r["accept_len"] == z[f"{dt}/accept_len"][b]
        assert r["next_token"] == z[f"{dt}/next_token"][b]
        n = r["accept_len"]
        assert r["tokens"].tolist() == z[f"{dt}/tokens"][b][:n].tolist()
        assert r["indices"].tolist() == z[f"{dt}/indices"][b][:n].tolist()


def test_verify_sequence_matches_reference():
    z = load("verify.npz")
    lg = torch.from_numpy(z["seq/logits_bits"]).view(torch.bfloat16)
    am = O.row_argmax(lg)
    for b in range(lg.shape[0]):
        r = O.verify_sequence(am[b], z["seq/tokens"][b])
        assert r["accept_len"] == z["seq/accept_len"][b]
        assert r["next_token"] == z["seq/next_token"][b]


def test_kv_compact_matches_reference():
    z = load("verify.npz")
    init = z["kv/init_bits"]                # [2L, 1, H, ML, DH] int16 bit patterns
    for c, (b, start) in enumerate(z["kv/cases"]):
        al = int(z["bf16/accept_len"][b])
        ind = z["bf16/indices"][b][:al]
        kv = [init[i, 0].copy() for i in range(init.shape[0])]
        new_len = O.kv_compact(kv, int(start), ind, al)
        assert new_len == start + al
        assert np.array_equal(np.stack(kv)[:, None], z["kv/after_bits"][c])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_decode_loop_matches_reference(name):
    z = load("decode_loop.npz")
    full = z[f"{name}/full"]
    plen = int(z[f"{name}/plen"])
    out, acc = O.generate_sam_only(full[:plen], full, 10 ** 9, 40, 4.0)
    ref = z[f"{name}/new_tokens"]
    assert out[:len(ref)] == ref.tolist()
    k = len(z[f"{name}/accepts"])
    assert acc[:k] == z[f"{name}/accepts"].tolist()


def test_heap_matches_cpython_heapq():
    import heapq
    rng = np.random.default_rng(5)
    for _ in range(200):
        mine, ref = O._Heap(), []
        serial = 0
        for _ in range(int(rng.integers(1, 80))):
            if rng.random() < 0.6 or not ref:
                key = float(rng.integers(0, 4)) / 4.0
                item = (key, serial)
                serial += 1
                mine.push(item)
                # heapq compares whole tuples; wrap so that only the key is compared
                heapq.heappush(ref, _KeyOnly(item))
            else:
                assert mine.pop() == heapq.heappop(ref).item


class _KeyOnly:
    def __init__(self, item):
        self.item = item

    def __lt__(self, other):
        return self.item[0] < other.item[0]


# --------------------------------------------------------------------------------------
# the C restatement (oracle/sam_oracle.c) against the same fixtures
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["v2", "v4", "mix1k", "mix4k", "light2k"])
def test_c_oracle_dyn_matches_reference(name):
    from c_oracle import CSam
    z = load("dyn_sam.npz")
    stream, cuts = z[f"{name}/stream"], z[f"{name}/cuts"]
    so = unragged(z[f"{name}/draft_so_flat"], z[f"{name}/draft_so_offs"])
    sam = CSam(len(stream) + 8)
    lo = 0
    for k, hi in enumerate(cuts):
        sam.extend(stream[lo:hi])
        lo = hi
        tok = int(stream[hi])
        i, l = sam.peek(tok)
        assert (i, l) == (z[f"{name}/index"][k], z[f"{name}/match"][k])
        assert sam.draft_samd(i, tok, 16) == z[f"{name}/draft16"][k].tolist()
        assert sam.draft_samd(i, tok, 40) == z[f"{name}/draft40"][k].tolist()
        assert sam.draft_so(i, l, tok, 40, 4.0) == so[k]
    link, length, end = sam.export()
    assert link.tolist() == z[f"{name}/link"].tolist()
    assert length.tolist() == z[f"{name}/length"].tolist()
    assert end.tolist() == z[f"{name}/min_endpos"].tolist()


def test_c_oracle_static_and_selection_match_reference():
    from c_oracle import CSam
    z = load("draft_select.npz")
    docs = docs_of(z)
    st = CSam.build(docs, 2)
    zs = load("static_sam.npz")
    assert st.export()[0].tolist() == zs["mid/link"].tolist()          # same corpus as static_sam.npz "mid"
    cuts_all = unragged(z["cuts_flat"], z["cuts_offs"])
    k = 0
    for r, stream in enumerate(z["streams"]):
        dyn = CSam(len(stream) + 8)
        st.reset_cursor()
        lo = 0
        for hi in cuts_all[r]:
            dyn.extend(stream[lo:hi])
            st.advance(stream[lo:hi])
            lo = hi
            kind, seq, _ = dyn.select_samd(st, int(stream[hi]), 16, 5, 5)
            assert kind == z["samd_source"][k]
            if kind != 2:
                assert seq == z["samd_seq"][k].tolist()
            k += 1


def test_token_recycle_oracle_matches_reference():
    """row_topk / recycle_update / recycle_gen_draft against the reference's TokenRecycle (tie-free rows)."""
    z = load("recycle.npz")
    tree = unragged(z["tree_flat"], z["tree_offs"])
    cache = {}
    for s in range(z["tokens"].shape[0]):
        logits = torch.from_numpy(z["logits_bits"][s].view(np.int16)).view(torch.bfloat16)
        topk = O.row_topk(logits)
        assert np.array_equal(topk, z["topk"][s])
        assert np.array_equal(topk[:, 0], O.row_argmax(logits))
        O.recycle_update(cache, z["tokens"][s], topk)
        for q, st in enumerate(z["starts"][s]):
            assert O.recycle_gen_draft(cache, tree, int(st)) == z["drafts"][s, q].tolist()
    assert sorted(cache) == z["cache_keys"].tolist()
    assert [cache[k] for k in sorted(cache)] == z["cache_vals"].tolist()


def test_row_topk_refinement_on_ties():
    """Equal values: lowest indices first; NaN above +inf; the selected VALUES equal torch.topk's."""
    x = torch.tensor([[1.0, 5.0, 5.0, float("nan"), 5.0, -0.0, 0.0, float("inf"), 2.0, 5.0, 1.0, 1.0]])
    got = O.row_topk(x, 8)[0].tolist()
    assert got == [3, 7, 1, 2, 4, 9, 8, 0]
    rng = np.random.default_rng(5)
    y = torch.from_numpy(rng.integers(0, 12, size=(50, 40)).astype(np.float32))
    idx = O.row_topk(y, 8)
    vals = np.take_along_axis(y.numpy(), idx, axis=1)
    assert np.array_equal(vals, y.topk(8).values.numpy())


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_typical_acceptance_oracle_matches_reference(ci):
    """samd/utils.py:142-184: the reference's eval_posterior (greedy=False) was run with random.random() replaced by a
    recorded stream (oracle/gen_sampling_golden.py); the oracle's restatement, fed the same stream, must take the same
    decisions and return the same next-token distribution - and consume the same number of draws."""
    z = load("sampling.npz")
    temp, top_p, top_k = z["configs"][ci]
    det, sps = z[f"c{ci}/det"], z[f"c{ci}/det_sample_p"]
    for (k0, n_draws, best, acc), sp in zip(det.tolist(), sps):
        k = [0]

        def draw():
            u = O.philox_uniform(5000 + ci, k0 + k[0])
            k[0] += 1
            return u

        got = O.verify_typical(z["logits"], z["tree_tokens"], z["retrieve"].astype(np.int64), float(temp), float(top_p), int(top_k), draw)
        assert (got["best"], got["accept_len"]) == (best, acc)
        assert k[0] - 1 == n_draws                       # (+1: the oracle also draws the next token)
        assert np.allclose(got["sample_p"], sp, rtol=1e-4, atol=1e-7)
