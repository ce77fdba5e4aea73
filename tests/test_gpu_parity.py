"""GPU parity tests: the CUDA path (through the C ABI, libsamd_b200.so) against the golden
fixtures produced by the reference's own classes, and against the CPU oracle on seeded
inputs.  Bit-exact: all results are integers / token ids / byte moves."""
import numpy as np
import pytest
import torch

import samd_oracle as O
from helpers import load, unragged, docs_of

pytestmark = pytest.mark.gpu

BIG = 1 << 20


def _engine_mod():
    from samd_b200 import engine, _cabi
    return engine, _cabi


def _dev_i32(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).cuda()


# --------------------------------------------------------------------------------------
def test_dyn_sam_golden():
    """DynSAM add_tokens / lookup / gen_draft (both flavours) on the reference's golden streams,
    all seven streams as one batch with ragged per-step counts."""
    E, K = _engine_mod()
    z = load("dyn_sam.npz")
    names = [str(n) for n in z["names"]]
    B = len(names)
    streams = [z[f"{n}/stream"] for n in names]
    cuts = [z[f"{n}/cuts"].tolist() for n in names]
    n_steps = max(len(c) for c in cuts)
    dyn = E.DynSamBatch(B, 8192)
    e16 = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=0, len_threshold=-BIG)
    e40 = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=40, len_bias=0, len_threshold=-BIG)
    eso = E.DraftEngine(dyn, None, K.FLAVOUR_SAM_ONLY, n_predicts=40, len_bias=0, alpha=4.0)
    so_ref = [unragged(z[f"{n}/draft_so_flat"], z[f"{n}/draft_so_offs"]) for n in names]
    for s in range(n_steps):
        lo = [0 if s == 0 else (cuts[r][s - 1] if s - 1 < len(cuts[r]) else None) for r in range(B)]
        hi = [cuts[r][s] if s < len(cuts[r]) else None for r in range(B)]
        width = max((h - l) for l, h in zip(lo, hi) if h is not None)
        tok = np.zeros((B, width), dtype=np.int32)
        cnt = np.zeros(B, dtype=np.int32)
        start = np.zeros(B, dtype=np.int32)
        for r in range(B):
            if hi[r] is None:
                continue
            cnt[r] = hi[r] - lo[r]
            tok[r, :cnt[r]] = streams[r][lo[r]:hi[r]]
            start[r] = streams[r][hi[r]]
        st = _dev_i32(start)
        e16.step(_dev_i32(tok), _dev_i32(cnt), st)
        e40.step(None, None, st)
        eso.step(None, None, st)
        torch.cuda.synchronize()
        for r, n in enumerate(names):
            if hi[r] is None:
                continue
            assert e16.index_dyn[r].item() == z[f"{n}/index"][s], (n, s)
            assert e16.match_dyn[r].item() == z[f"{n}/match"][s], (n, s)
            assert e16.out_type[r].item() == K.DRAFT_DYN_SEQ
            assert e16.draft[r].tolist() == z[f"{n}/draft16"][s].tolist(), (n, s)
            assert e40.draft[r].tolist() == z[f"{n}/draft40"][s].tolist(), (n, s)
            k = eso.draft_len[r].item()
            assert eso.draft[r, :k].tolist() == so_ref[r][s], (n, s)
    for r, n in enumerate(names):
        ex = dyn.export(r)
        assert ex["overflow"] == 0
        assert np.array_equal(ex["link"], z[f"{n}/link"])
        assert np.array_equal(ex["length"], z[f"{n}/length"])
        assert np.array_equal(ex["min_endpos"], z[f"{n}/min_endpos"])
        assert [ex["cur_index"], ex["cur_length"]] == z[f"{n}/cursor"].tolist()
        assert np.array_equal(ex["text"][1:], streams[r][:cuts[r][-1]])


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny", "small", "mid"])
def test_static_sam_golden(name, tmp_path):
    """StaticSAM build (state numbering, counts, top-k), transfer_tokens, lookup, gen_draft
    (sequence + sam_only tree) against the reference's golden outputs; save/load round trip."""
    E, K = _engine_mod()
    z = load("static_sam.npz")
    docs = docs_of(z, name)
    st = E.StaticSamDevice.build(docs, int(z[f"{name}/eos"]), with_counts=True)
    ex = st.export()
    assert np.array_equal(ex["link"], z[f"{name}/link"])
    assert np.array_equal(ex["length"], z[f"{name}/length"])
    assert np.array_equal(ex["min_endpos"], z[f"{name}/min_endpos"])
    assert np.array_equal(ex["cnt_endpos"], z[f"{name}/cnt_endpos"])
    assert np.array_equal(ex["topk"][:, :, 0], z[f"{name}/topk_tok"])
    assert np.array_equal(ex["topk"][:, :, 1], z[f"{name}/topk_idx"])
    path = str(tmp_path / "sam.bin")
    st.save(path)
    st2 = E.StaticSamDevice.load(path)
    assert (st2.n_states, st2.n_edges, st2.n_tokens) == (st.n_states, st.n_edges, st.n_tokens)

    q = z[f"{name}/queries"]
    steps = z[f"{name}/steps"]
    nq = q.shape[0]
    per_q = [[] for _ in range(nq)]
    for k, (qi, p) in enumerate(steps):
        per_q[qi].append((int(p), k))
    trees = unragged(z[f"{name}/tree_tok_flat"], z[f"{name}/tree_offs"])
    rets = unragged(z[f"{name}/tree_ret_flat"], z[f"{name}/tree_ret_offs"])
    for sam in (st, st2):
        dyn = E.DynSamBatch(nq, 256)
        # len_bias = -BIG makes the static automaton win every selection -> static sequence draft
        eng = E.DraftEngine(dyn, sam, K.FLAVOUR_SAMD, n_predicts=16, len_bias=-BIG, len_threshold=-BIG)
        pos = [0] * nq
        for s in range(max(len(x) for x in per_q)):
            tok = np.zeros((nq, 8), dtype=np.int32)
            cnt = np.zeros(nq, dtype=np.int32)
            start = np.zeros(nq, dtype=np.int32)
            live = []
            for qi in range(nq):
                if s >= len(per_q[qi]):
                    continue
                p, k = per_q[qi][s]
                cnt[qi] = p - pos[qi]
                tok[qi, :cnt[qi]] = q[qi, pos[qi]:p]
                start[qi] = q[qi, p]
                pos[qi] = p
                live.append((qi, k))
            st_dev = _dev_i32(start)
            eng.step(_dev_i32(tok), _dev_i32(cnt), st_dev)
            # tree draft with the fixture's budget rule: match = max(l - 2, 0), bias 0
            m = torch.clamp(eng.match_static - 2, min=0)
            n, mp = 40, 40
            mk = lambda *sh: torch.zeros(*sh, dtype=torch.int32, device="cuda")
            t_tok, t_par, t_dep, t_n, t_shape, t_ret = mk(nq, n), mk(nq, n), mk(nq, n), mk(nq), mk(nq, 2), mk(nq, mp, n)
            K.check(K.lib().samd_static_tree_draft(
                sam.handle, nq, None, eng.index_static.data_ptr(), m.data_ptr(), st_dev.data_ptr(), n, 4.0, 8, 0,
                t_tok.data_ptr(), t_par.data_ptr(), t_dep.data_ptr(), t_n.data_ptr(), t_ret.data_ptr(), mp, n,
                t_shape.data_ptr(), K.stream_ptr()))
            torch.cuda.synchronize()
            for qi, k in live:
                assert eng.index_static[qi].item() == z[f"{name}/index"][k]
                assert eng.match_static[qi].item() == z[f"{name}/match"][k]
                assert eng.out_type[qi].item() == K.DRAFT_STATIC_SEQ
                assert eng.draft[qi].tolist() == z[f"{name}/draft16"][k].tolist()
                nn = t_n[qi].item()
                assert t_tok[qi, :nn].tolist() == trees[k]
                off = z[f"{name}/tree_offs"][k]
                assert t_dep[qi, :nn].tolist() == z[f"{name}/tree_depth_flat"][off:off + nn].tolist()
                shape = tuple(t_shape[qi].tolist())
                assert shape == tuple(z[f"{name}/tree_ret_shape"][k])
                assert t_ret[qi, :shape[0], :shape[1]].reshape(-1).tolist() == rets[k]


# --------------------------------------------------------------------------------------
def test_draft_selection_golden():
    """DraftModel.lookup / update of both packages, six requests as one batch."""
    E, K = _engine_mod()
    z = load("draft_select.npz")
    docs = docs_of(z)
    st = E.StaticSamDevice.build(docs, 2, with_counts=True)
    streams = z["streams"]
    B = streams.shape[0]
    cuts = unragged(z["cuts_flat"], z["cuts_offs"])
    so_tok = unragged(z["so_tok_flat"], z["so_tok_offs"])
    so_ret = unragged(z["so_ret_flat"], z["so_ret_offs"])
    base = np.cumsum([0] + [len(c) for c in cuts])
    dyn_a, dyn_b = E.DynSamBatch(B, 1024), E.DynSamBatch(B, 1024)
    ea = E.DraftEngine(dyn_a, st, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
    eb = E.DraftEngine(dyn_b, st, K.FLAVOUR_SAM_ONLY, n_predicts=40, len_bias=5, alpha=4.0)
    n_steps = max(len(c) for c in cuts)
    for s in range(n_steps):
        width = max(cuts[r][s] - (cuts[r][s - 1] if s else 0) for r in range(B) if s < len(cuts[r]))
        tok = np.zeros((B, width), dtype=np.int32)
        cnt = np.zeros(B, dtype=np.int32)
        start = np.zeros(B, dtype=np.int32)
        for r in range(B):
            if s >= len(cuts[r]):
                continue
            lo, hi = (cuts[r][s - 1] if s else 0), cuts[r][s]
            cnt[r] = hi - lo
            tok[r, :cnt[r]] = streams[r][lo:hi]
            start[r] = streams[r][hi]
        t, c, sd = _dev_i32(tok), _dev_i32(cnt), _dev_i32(start)
        ea.step(t, c, sd)
        eb.step(t, c, sd)
        eb.tree_draft(sd)
        torch.cuda.synchronize()
        for r in range(B):
            if s >= len(cuts[r]):
                continue
            k = base[r] + s
            src = z["samd_source"][k]
            assert ea.out_type[r].item() == {0: K.DRAFT_DYN_SEQ, 1: K.DRAFT_STATIC_SEQ, 2: K.DRAFT_TREE_MODEL}[int(src)]
            if src != 2:
                assert ea.draft[r].tolist() == z["samd_seq"][k].tolist()
            if z["so_type"][k] == 0:
                assert eb.out_type[r].item() == K.DRAFT_DYN_SEQ
                assert eb.draft[r, :eb.draft_len[r].item()].tolist() == so_tok[k]
            else:
                assert eb.out_type[r].item() == K.DRAFT_STATIC_TREE
                nn = eb.tree_n[r].item()
                assert eb.tree_tokens[r, :nn].tolist() == so_tok[k]
                shp = eb.tree_shape[r].tolist()
                assert eb.tree_retrieve[r, :shp[0], :shp[1]].reshape(-1).tolist() == so_ret[k]


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_verify_golden(dt):
    """gather + eval_posterior + update_state slices against the reference (ties, NaN, +-0, pads)."""
    E, K = _engine_mod()
    z = load("verify.npz")
    bits = torch.from_numpy(z["bf16/logits_bits"]).view(torch.bfloat16)
    lg = (bits if dt == "bf16" else bits.float().to(torch.float16)).cuda()
    B, T, V = lg.shape
    ri = _dev_i32(z["retrieve"])
    ver = E.Verifier(B, T)
    cache_len = torch.full((B,), 100, dtype=torch.int32, device="cuda")
    out = ver.verify(lg, _dev_i32(z["tree_tokens"]), ri, cache_len=cache_len, want_argmax=True)
    # second launch on the same scratch must give the same answer (self re-arming counters)
    out2 = ver.verify(lg, _dev_i32(z["tree_tokens"]), ri, cache_len=None, want_argmax=True)
    torch.cuda.synchronize()
    for o in (out, out2):
        assert np.array_equal(o["node_argmax"].cpu().numpy(), z[f"{dt}/node_argmax"])
        assert np.array_equal(o["best"].cpu().numpy(), z[f"{dt}/best"])
        assert np.array_equal(o["accept_len"].cpu().numpy(), z[f"{dt}/accept_len"])
        assert np.array_equal(o["next_token"].cpu().numpy(), z[f"{dt}/next_token"])
        tk, ix = o["tokens"].cpu().numpy(), o["indices"].cpu().numpy()
        for b in range(B):
            n = z[f"{dt}/accept_len"][b]
            assert tk[b, :n].tolist() == z[f"{dt}/tokens"][b][:n].tolist()
            assert ix[b, :n].tolist() == z[f"{dt}/indices"][b][:n].tolist()
    assert np.array_equal(cache_len.cpu().numpy(), 100 + z[f"{dt}/accept_len"])


def test_verify_sequence_golden():
    E, K = _engine_mod()
    z = load("verify.npz")
    lg = torch.from_numpy(z["seq/logits_bits"]).view(torch.bfloat16).cuda()
    B, T, V = lg.shape
    ver = E.Verifier(B, T)
    out = ver.verify(lg, _dev_i32(z["seq/tokens"]), None)
    torch.cuda.synchronize()
    assert np.array_equal(out["accept_len"].cpu().numpy(), z["seq/accept_len"])
    assert np.array_equal(out["next_token"].cpu().numpy(), z["seq/next_token"])
    assert (out["best"] == 0).all()


@pytest.mark.parametrize("overlap", [0, 1, 2, 3])
def test_kv_compaction_golden(overlap):
    """select_indices through the fused kernel: the fixture's 16 cases as one batch of 16 requests.  All three flows of
    the kernel: 0 = stream / barrier / walks / barrier / row moves (the default), 1 = walks as requests complete and
    ticketed row moves, 2 = early walks + L2 prefetch of the source rows, 3 = early walks only (the measured alternatives,
    DESIGN.md)."""
    E, K = _engine_mod()
    K.lib().samd_verify_set_overlap(overlap)
    try:
        _kv_compaction_golden(E, K)
    finally:
        K.lib().samd_verify_set_overlap(0)


def _kv_compaction_golden(E, K):
    z = load("verify.npz")
    init = torch.from_numpy(z["kv/init_bits"]).view(torch.bfloat16)            # [2L, 1, H, ML, DH]
    cases = z["kv/cases"]
    n_case = len(cases)
    bits = torch.from_numpy(z["bf16/logits_bits"]).view(torch.bfloat16)
    sel = [int(b) for b, _ in cases]
    lg = bits[sel].contiguous().cuda()
    tok = _dev_i32(z["tree_tokens"][sel])
    kv = [init[i].repeat(n_case, 1, 1, 1).contiguous().cuda() for i in range(init.shape[0])]
    cache_len = _dev_i32(np.array([s for _, s in cases]))
    ver = E.Verifier(n_case, lg.shape[1])
    ver.bind_kv(kv)
    out = ver.verify(lg, tok, _dev_i32(z["retrieve"]), cache_len=cache_len, move_kv=True)
    torch.cuda.synchronize()
    after = z["kv/after_bits"]                                                 # [case, 2L, 1, H, ML, DH]
    for c in range(n_case):
        got = torch.stack([t[c] for t in kv]).view(torch.int16).cpu().numpy()
        assert np.array_equal(got, after[c][:, 0]), c
        assert cache_len[c].item() == cases[c][1] + z["bf16/accept_len"][sel[c]]
    assert np.array_equal(out["accept_len"].cpu().numpy(), z["bf16/accept_len"][sel])
    # a second launch on the same scratch (self re-arming counters): nothing left to move, lengths bump again
    before = [t.clone() for t in kv]
    ver.verify(lg, tok, _dev_i32(z["retrieve"]), cache_len=None, move_kv=False)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(before, kv))


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_decode_loop_golden(name):
    """prefill update -> {lookup -> fake-LM logits -> fused verify -> update}* entirely through
    the C ABI; token stream and accept lengths must equal the reference's loop."""
    E, K = _engine_mod()
    z = load("decode_loop.npz")
    full = z[f"{name}/full"]
    plen = int(z[f"{name}/plen"])
    vocab = int(z["vocab"])
    ref_tokens = z[f"{name}/new_tokens"].tolist()
    ref_acc = z[f"{name}/accepts"].tolist()
    dyn = E.DynSamBatch(1, 4096)
    eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAM_ONLY, n_predicts=40, len_bias=5, alpha=4.0)
    ver = E.Verifier(1, 40)
    eng.step(_dev_i32(full[None, :plen]), None, None)
    pos = plen
    start = _dev_i32([full[pos]])
    out_tokens, accepts = [], []
    for _ in range(len(ref_acc)):
        eng.step(None, None, start)
        n = eng.draft_len[0].item()
        logits = torch.zeros(1, 40, vocab, dtype=torch.bfloat16, device="cuda")
        idx = torch.as_tensor(full[pos + 1:pos + 1 + n].astype(np.int64)).cuda()
        logits[0, torch.arange(n, device="cuda"), idx] = 1.0
        res = ver.verify(logits, eng.draft, None, n_nodes=eng.draft_len)
        al = res["accept_len"][0].item()
        acc = res["tokens"][0, :al].tolist()
        eng.step(res["tokens"], res["accept_len"], None)          # device-to-device, no host round trip needed
        start = res["next_token"]
        out_tokens.extend(acc)
        accepts.append(al)
        pos += al
    assert accepts == ref_acc
    assert out_tokens[:len(ref_tokens)] == ref_tokens


# --------------------------------------------------------------------------------------
def test_dyn_batch_vs_oracle_8k():
    """Config-2 shape at reduced batch: 32 requests x 8192-token prompts, 24 decode steps of 1-8
    appended tokens, both flavours, against the CPU oracle (clone-heavy and clone-light streams)."""
    E, K = _engine_mod()
    from samd_b200 import synth
    B, N, steps = 32, 8192, 24
    total = N + 8 * steps + 1
    streams = [synth.copy_mix(total, 32000, 2000 + r, uniform_fresh=(r % 2 == 1)) for r in range(B)]
    rng = np.random.default_rng(9)
    dyn = E.DynSamBatch(B, N + 8 * steps + 8)
    e16 = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
    eso = E.DraftEngine(dyn, None, K.FLAVOUR_SAM_ONLY, n_predicts=40, len_bias=5, alpha=4.0)
    oracles = [O.Automaton() for _ in range(B)]
    prompt = np.stack([s[:N] for s in streams]).astype(np.int32)
    e16.step(_dev_i32(prompt), None, None)
    for r in range(B):
        oracles[r].extend(streams[r][:N])
    pos = [N] * B
    for s in range(steps):
        cnt = rng.integers(1, 9, size=B).astype(np.int32)
        tok = np.zeros((B, 8), dtype=np.int32)
        start = np.zeros(B, dtype=np.int32)
        for r in range(B):
            tok[r, :cnt[r]] = streams[r][pos[r]:pos[r] + cnt[r]]
            oracles[r].extend(streams[r][pos[r]:pos[r] + cnt[r]])
            pos[r] += cnt[r]
            start[r] = streams[r][pos[r]]
        sd = _dev_i32(start)
        e16.step(_dev_i32(tok), _dev_i32(cnt), sd)
        eso.step(None, None, sd)
        torch.cuda.synchronize()
        for r in range(B):
            typ, seq, info = O.select_samd(oracles[r], None, int(start[r]), 16, 5, 5)
            assert e16.match_dyn[r].item() == info["match_dyn"] and e16.index_dyn[r].item() == info["index_dyn"]
            if typ == "sequence":
                assert e16.out_type[r].item() == K.DRAFT_DYN_SEQ
                assert e16.draft[r].tolist() == seq
            else:
                assert e16.out_type[r].item() == K.DRAFT_TREE_MODEL
            want = O.dyn_draft_sam_only(oracles[r], info["index_dyn"], info["match_dyn"], int(start[r]), 40, 4.0)
            assert eso.draft[r, :eso.draft_len[r].item()].tolist() == want
    for r in (0, 1, B - 1):
        ex = dyn.export(r, with_text=False)
        assert ex["n_states"] == oracles[r].n_states and ex["n_edges"] == oracles[r].n_edges
        assert np.array_equal(ex["link"], np.array(oracles[r].link))
        assert np.array_equal(ex["min_endpos"], np.array(oracles[r].first_end))


@pytest.mark.parametrize("tma", [0, 1])
def test_verify_full_size_properties(tma):
    """Config-4 shape (B=64, T=61, V=32000, bf16): node argmax equals torch.argmax, outputs equal
    the oracle walk, and the compacted KV rows equal an index_select reference.  tma = 1: the logits stream staged
    through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier ring; a measured alternative, default off)."""
    E, K = _engine_mod()
    K.lib().samd_verify_set_tma(tma)
    try:
        _verify_full_size(E, K)
    finally:
        K.lib().samd_verify_set_tma(0)


def _verify_full_size(E, K):
    from samd_b200 import synth
    B, T, V = 64, 61, 32000
    ri_np = synth.tree_retrieve_indices(synth.token_recycle_tree())
    rng = np.random.default_rng(4000)
    tree_tokens = rng.integers(3, V, size=(B, T)).astype(np.int32)
    logits, _ = synth.planted_logits(B, T, V, tree_tokens, ri_np, seed=4000, device="cuda")
    L, H, ML, DH = 4, 8, 256, 128
    kv = [torch.randn(B, H, ML, DH, device="cuda").to(torch.bfloat16) for _ in range(2 * L)]
    kv_ref = [t.clone() for t in kv]
    cache_len = torch.randint(16, 150, (B,), dtype=torch.int32, device="cuda")
    start = cache_len.clone()
    ver = E.Verifier(B, T)
    ver.bind_kv(kv)
    out = ver.verify(logits, _dev_i32(tree_tokens), _dev_i32(ri_np), cache_len=cache_len, want_argmax=True)
    torch.cuda.synchronize()
    am = torch.argmax(logits, dim=-1)
    assert torch.equal(out["node_argmax"].long(), am)
    am_np = am.cpu().numpy()
    for b in range(B):
        r = O.verify_greedy(am_np[b], tree_tokens[b], ri_np.astype(np.int64))
        al = r["accept_len"]
        assert out["best"][b].item() == r["best"] and out["accept_len"][b].item() == al
        assert out["next_token"][b].item() == r["next_token"]
        assert out["tokens"][b, :al].tolist() == r["tokens"].tolist()
        assert out["indices"][b, :al].tolist() == r["indices"].tolist()
        s0 = start[b].item()
        src = torch.as_tensor(s0 + r["indices"], device="cuda")
        for t, t_ref in zip(kv, kv_ref):
            want = t_ref[b].clone()
            want[:, s0:s0 + al] = t_ref[b].index_select(1, src)
            assert torch.equal(t[b], want)
        assert cache_len[b].item() == s0 + al


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_document_sharded_static_sam(world):
    """per-shard packed keys + max-reduce + draft from the replicated corpus.  All shards
    are built on this one GPU and the all-reduce is replaced by torch.maximum (the NCCL/gloo reduction
    itself is covered by tests/test_dist_gloo.py and bench.py --gpus N)."""
    E, K = _engine_mod()
    from samd_b200 import dist as D, synth
    docs = [d.tolist() for d in synth.make_corpus(30000, 500, 61, doc_len=(16, 96), singletons=True)]
    eos = synth.EOS
    nq, qlen = 64, 40
    q = synth.corpus_queries([np.array(d) for d in docs], nq, qlen, 500, 62)
    q[q == eos] = 9                                    # EOS-free queries: equality with the global automaton is exact
    shards = [D.ShardedStaticSam(docs, eos, g, world, nq, torch.device("cuda")) for g in range(world)]
    whole = O.build_static(docs, eos)
    assert sum(s.sam.n_tokens for s in shards) == whole.n and shards[-1].offset + shards[-1].sam.n_tokens == whole.n
    pos = 0
    rng = np.random.default_rng(63)
    cursors = [[0, 0] for _ in range(nq)]
    while pos < qlen - 1:
        k = int(rng.integers(1, 9))
        k = min(k, qlen - 1 - pos)
        tok = _dev_i32(q[:, pos:pos + k])
        start = _dev_i32(q[:, pos + k])
        keys = None
        for s in shards:
            s.advance(tok)
            kk = s.local_keys(start).clone()
            keys = kk if keys is None else torch.maximum(keys, kk)
        match, draft = shards[0].draft(keys, start, 16)
        torch.cuda.synchronize()
        pos += k
        for i in range(nq):
            st, ln = cursors[i]
            for t in q[i, pos - k:pos]:
                st, ln = whole.step(st, ln, int(t))
            cursors[i] = [st, ln]
            s2, l2 = whole.step(st, ln, int(q[i, pos]))
            assert match[i].item() == l2
            if l2:
                assert draft[i].tolist() == O.static_draft_samd(whole, s2, int(q[i, pos]), 16)


# --------------------------------------------------------------------------------------
def test_c2_full_size_against_c_oracle():
    """BASELINE config 2 at FULL size: 1024 requests x 8192-token prompts, then 6 decode steps of 1-8
    appended tokens.  Every request's match length, state index and 16-token draft is compared with the
    oracle's C restatement (oracle/sam_oracle.c), and the automaton sizes (states, edges, clones) of all
    1024 arenas must add up to the oracle's."""
    E, K = _engine_mod()
    from c_oracle import CSam
    from samd_b200 import synth
    R, N, steps = 1024, 8192, 6
    total = N + 8 * steps + 1
    rng = np.random.default_rng(20)
    streams = np.stack([synth.copy_mix(total, 32000, 9000 + r, uniform_fresh=(r % 4 == 3)).astype(np.int32) for r in range(R)])
    dyn = E.DynSamBatch(R, total + 8)
    eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=-BIG)
    eng.step(_dev_i32(streams[:, :N]), None, None)
    oracles = [CSam(total + 8) for _ in range(R)]
    for r in range(R):
        oracles[r].extend(streams[r, :N])
    pos = np.full(R, N)
    for s in range(steps):
        cnt = rng.integers(1, 9, size=R).astype(np.int32)
        tok = np.zeros((R, 8), dtype=np.int32)
        for r in range(R):
            tok[r, :cnt[r]] = streams[r, pos[r]:pos[r] + cnt[r]]
            oracles[r].extend(tok[r, :cnt[r]])
        pos += cnt
        start = streams[np.arange(R), pos]
        eng.step(_dev_i32(tok), _dev_i32(cnt), _dev_i32(start))
        torch.cuda.synchronize()
        idx, mlen, draft = eng.index_dyn.cpu().numpy(), eng.match_dyn.cpu().numpy(), eng.draft.cpu().numpy()
        for r in range(R):
            kind, seq, info = oracles[r].select_samd(None, int(start[r]), 16, 5, -BIG)
            assert kind == 0 and (info[0], info[1]) == (idx[r], mlen[r]), (s, r)
            assert seq == draft[r].tolist(), (s, r)
    st = dyn.stats()
    infos = [o.info() for o in oracles]
    assert st["overflowed"] == 0
    assert st["n_states"] == sum(i["n_states"] for i in infos)
    assert st["n_edges"] == sum(i["n_edges"] for i in infos)
    assert st["n_clones"] == sum(i["n_clones"] for i in infos)
    assert st["tokens"] == int(pos.sum())


def test_verify_compact_is_graph_replayable():
    """The fused kernel keeps its launch epoch in device memory, so a captured launch can be replayed: three
    replays on fresh KV copies must each produce the reference result (golden KV fixture, 16 requests)."""
    E, K = _engine_mod()
    z = load("verify.npz")
    init = torch.from_numpy(z["kv/init_bits"]).view(torch.bfloat16)
    cases = z["kv/cases"]
    n_case = len(cases)
    sel = [int(b) for b, _ in cases]
    lg = torch.from_numpy(z["bf16/logits_bits"]).view(torch.bfloat16)[sel].contiguous().cuda()
    tok = _dev_i32(z["tree_tokens"][sel])
    ri = _dev_i32(z["retrieve"])
    kv0 = [init[i].repeat(n_case, 1, 1, 1).contiguous().cuda() for i in range(init.shape[0])]
    kv = [t.clone() for t in kv0]
    start = _dev_i32(np.array([s for _, s in cases]))
    cache_len = start.clone()
    ver = E.Verifier(n_case, lg.shape[1])
    ver.bind_kv(kv)
    out = ver.verify(lg, tok, ri, cache_len=cache_len, move_kv=True)          # eager warm-up allocates `out`
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ver.verify(lg, tok, ri, cache_len=cache_len, move_kv=True, out=out)
    for _ in range(3):
        for t, t0 in zip(kv, kv0):
            t.copy_(t0)
        cache_len.copy_(start)
        out["accept_len"].zero_()
        g.replay()
        torch.cuda.synchronize()
        for c in range(n_case):
            got = torch.stack([t[c] for t in kv]).view(torch.int16).cpu().numpy()
            assert np.array_equal(got, z["kv/after_bits"][c][:, 0]), c
        assert np.array_equal(out["accept_len"].cpu().numpy(), z["bf16/accept_len"][sel])
        assert np.array_equal(cache_len.cpu().numpy(), start.cpu().numpy() + z["bf16/accept_len"][sel])


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,V,P,D,dt", [
    (1, 1, 5, 1, 1, "bf16"), (3, 7, 1001, 5, 4, "fp16"), (5, 61, 32001, 30, 6, "bf16"), (130, 3, 4099, 3, 3, "bf16"),
    (9, 40, 2500, 40, 10, "fp32"), (17, 12, 777, 33, 8, "bf16"), (700, 2, 64, 2, 2, "fp16"), (2, 64, 151936, 12, 7, "bf16"),
])
def test_verify_random_shapes(B, T, V, P, D, dt):
    """Odd shapes through every code path of the verify kernel: vocabularies that are not a multiple of the vector
    width (padded row stride -> vector path with a ragged end; unpadded -> scalar path), more than 32 paths or more
    than 8 levels (the general walk), per-request path tables with -1 padding, ragged node counts, planted ties,
    and KV row moves - against torch.argmax, the oracle walk and an index_select reference."""
    E, K = _engine_mod()
    rng = np.random.default_rng(B * 1000 + T)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dt]
    for padded in (True, False):
        Vp = (V + 7) // 8 * 8 if padded else V
        store = torch.randn(B, T, Vp, device="cuda").to(tdt)
        logits = store[:, :, :V]
        # ties: copy the row maximum to a later column in a third of the rows (lowest index must win)
        am0 = logits.float().argmax(-1)
        if V > 2:
            tie_col = torch.clamp(am0 + 1 + torch.randint(0, V, am0.shape, device="cuda") % (V - 1), max=V - 1)
            mask = torch.rand(B, T, device="cuda") < 0.33
            vals = logits.gather(2, am0[..., None])
            logits.scatter_(2, tie_col[..., None], torch.where(mask[..., None], vals, logits.gather(2, tie_col[..., None])))
        n_nodes = rng.integers(1, T + 1, size=B).astype(np.int32)
        am = logits.float().argmax(-1).cpu().numpy()      # torch.argmax: first maximal index
        retrieve = np.full((B, P, D), -1, dtype=np.int32)
        n_paths = rng.integers(1, P + 1, size=B).astype(np.int32)
        tree_tokens = rng.integers(1, V, size=(B, T)).astype(np.int32)
        for b in range(B):
            for p in range(n_paths[b]):
                depth = int(rng.integers(1, D + 1))
                rest = np.sort(rng.choice(np.arange(1, n_nodes[b]), size=min(depth - 1, n_nodes[b] - 1), replace=False)) \
                    if n_nodes[b] > 1 else np.zeros(0, dtype=np.int64)
                path = np.concatenate([[0], rest]).astype(np.int32)
                retrieve[b, p, :len(path)] = path
                for j in range(1, len(path)):              # plant acceptances along the path
                    if rng.random() < 0.6:
                        tree_tokens[b, path[j]] = am[b, path[j - 1]]
        L, H, ML, DH = 2, 2, T + 40, 16
        kv = [torch.randn(B, H, ML, DH, device="cuda").to(torch.bfloat16) for _ in range(2 * L)]
        kv_ref = [t.clone() for t in kv]
        start = rng.integers(0, 30, size=B).astype(np.int32)
        cache_len = _dev_i32(start)
        ver = E.Verifier(B, T)
        ver.bind_kv(kv)
        out = None
        for rep in range(2):                               # the second launch runs on re-armed scratch
            for t, t0 in zip(kv, kv_ref):
                t.copy_(t0)
            cache_len.copy_(_dev_i32(start))
            out = ver.verify(logits, _dev_i32(tree_tokens), _dev_i32(retrieve), cache_len=cache_len, n_nodes=_dev_i32(n_nodes),
                             n_paths=_dev_i32(n_paths), want_argmax=True, out=out)
            torch.cuda.synchronize()
            got_am = out["node_argmax"].cpu().numpy()
            o = {k: v.cpu().numpy() for k, v in out.items()}
            for b in range(B):
                nr = n_nodes[b]
                assert np.array_equal(got_am[b, :nr], am[b, :nr]), (padded, b)
                r = O.verify_greedy(am[b, :nr], tree_tokens[b, :nr], retrieve[b, :n_paths[b]].astype(np.int64))
                al = r["accept_len"]
                assert o["best"][b] == r["best"] and o["accept_len"][b] == al and o["next_token"][b] == r["next_token"]
                assert o["tokens"][b, :al].tolist() == r["tokens"].tolist()
                assert o["indices"][b, :al].tolist() == r["indices"].tolist()
                if B <= 20:
                    src = torch.as_tensor(int(start[b]) + r["indices"], device="cuda")
                    for t, t_ref in zip(kv, kv_ref):
                        want = t_ref[b].clone()
                        want[:, start[b]:start[b] + al] = t_ref[b].index_select(1, src)
                        assert torch.equal(t[b], want), (padded, b)
            assert np.array_equal(cache_len.cpu().numpy(), start + o["accept_len"])
        ver.close()


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["zero_copy", "stage_in", "stage_both", "copy_engine"])
def test_step_host_matches_device_step(mode):
    """The host-buffer entry (DraftEngine.step_host: pinned buffers, graph-replayed; zero-copy, staged by copy kernels or
    by the copy engine) produces exactly what the device-buffer step produces, step after step."""
    E, K = _engine_mod()
    from samd_b200 import synth
    B, N, steps = 48, 512, 12
    streams = [synth.copy_mix(N + 8 * steps + 1, 500, 7000 + r) for r in range(B)]
    prompt = _dev_i32(np.stack([s[:N] for s in streams]))
    engines = []
    for _ in range(2):
        dyn = E.DynSamBatch(B, N + 8 * steps + 8)
        eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
        eng.step(prompt, None, None)
        engines.append(eng)
    dev_eng, host_eng = engines
    inp, res = host_eng.host_buffers(8)
    rng = np.random.default_rng(11)
    pos = [N] * B
    for s in range(steps):
        cnt = rng.integers(0, 9, size=B).astype(np.int32)
        tok = np.zeros((B, 8), dtype=np.int32)
        start = np.zeros(B, dtype=np.int32)
        for r in range(B):
            tok[r, :cnt[r]] = streams[r][pos[r]:pos[r] + cnt[r]]
            pos[r] += cnt[r]
            start[r] = streams[r][pos[r]]
        dev_eng.step(_dev_i32(tok), _dev_i32(cnt), _dev_i32(start))
        inp[:B] = torch.as_tensor(cnt)
        inp[B:2 * B] = torch.as_tensor(start)
        inp[2 * B:] = torch.as_tensor(tok).reshape(-1)
        host_eng.step_host(inp, res, mode=mode)
        want = dev_eng.out_buf.cpu().numpy()
        got = res.numpy()
        # type, match lengths, state indices, draft length: exact; draft tokens: up to draft_len
        assert np.array_equal(got[:6 * B], want[:6 * B]), s
        dl = want[5 * B:6 * B]
        gd, wd = got[6 * B:].reshape(B, -1), want[6 * B:].reshape(B, -1)
        for r in range(B):
            assert np.array_equal(gd[r, :dl[r]], wd[r, :dl[r]]), (s, r)


# --------------------------------------------------------------------------------------
def test_token_recycle_golden():
    """TokenRecycle.update fused into the verify launch + gen_draft, against the reference's own class
    (tests/golden/recycle.npz: tie-free rows, tokens repeating inside and across steps, starts without an entry)."""
    E, K = _engine_mod()
    z = load("recycle.npz")
    tree = unragged(z["tree_flat"], z["tree_offs"])
    steps, T, V = z["logits_bits"].shape
    table = E.RecycleTable(tree)
    ver = E.Verifier(1, T)
    from samd_b200 import synth
    ri = _dev_i32(synth.tree_retrieve_indices(tree))
    for s in range(steps):
        lg = torch.from_numpy(z["logits_bits"][s].view(np.int16)).view(torch.bfloat16).cuda().view(1, T, V)
        out = ver.verify(lg, _dev_i32(z["tokens"][s][None]), ri, recycle=table, want_argmax=True)
        torch.cuda.synchronize()
        assert np.array_equal(out["topk"][0].cpu().numpy(), z["topk"][s])
        assert np.array_equal(out["node_argmax"][0].cpu().numpy(), z["topk"][s][:, 0])
        got = table.gen_tree(_dev_i32(z["starts"][s])).cpu().numpy()
        assert np.array_equal(got, z["drafts"][s])
    d = table.as_dict()
    assert sorted(d) == z["cache_keys"].tolist()
    assert [d[k] for k in sorted(d)] == z["cache_vals"].tolist()
    # the stand-alone entry (rows as one-node requests) builds the same table from the same rows
    t2 = E.RecycleTable(tree)
    for s in range(steps):
        lg = torch.from_numpy(z["logits_bits"][s].view(np.int16)).view(torch.bfloat16).cuda()
        topk = t2.update(_dev_i32(z["tokens"][s]), lg)
        assert np.array_equal(topk.cpu().numpy(), z["topk"][s])
    assert t2.as_dict() == d


@pytest.mark.parametrize("B,T,V,dt", [(2, 5, 16, "bf16"), (3, 7, 1001, "fp16"), (4, 9, 32001, "bf16"), (64, 3, 4099, "bf16"),
                                      (5, 6, 2500, "fp32"), (1, 61, 151936, "bf16"), (300, 2, 64, "fp16")])
def test_verify_topk_random(B, T, V, dt):
    """Top-8 per row with heavy ties, NaNs, -inf floods and ragged trees, every load path (vector / scalar, chunked
    rows): indices must equal the oracle's (value descending, index ascending) refinement, the selected values must
    equal torch.topk's, and the recycle table must equal the reference's zip-order overwrite."""
    E, K = _engine_mod()
    rng = np.random.default_rng(V + B)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dt]
    for padded in (True, False):
        Vp = (V + 7) // 8 * 8 if padded else V
        # coarse values: about 40 distinct levels -> ties everywhere, also inside the top 8
        store = (torch.randint(-20, 20, (B, T, Vp), device="cuda").float() * 0.5).to(tdt)
        logits = store[:, :, :V]
        # -inf floods (fewer than 8 finite values in some rows), NaNs, +inf
        for _ in range(max(1, B * T // 3)):
            b, t = int(rng.integers(B)), int(rng.integers(T))
            kind = int(rng.integers(4))
            if kind == 0:
                logits[b, t, :] = float("-inf")
                keep = rng.choice(V, size=int(rng.integers(0, 7)), replace=False)
                logits[b, t, torch.as_tensor(keep, device="cuda", dtype=torch.long)] = 1.5
            elif kind == 1:
                logits[b, t, torch.as_tensor(rng.choice(V, size=3, replace=False), device="cuda")] = float("nan")
            elif kind == 2:
                logits[b, t, int(rng.integers(V))] = float("inf")
        n_nodes = rng.integers(1, T + 1, size=B).astype(np.int32)
        tokens = rng.integers(0, min(V, 40), size=(B, T)).astype(np.int32)       # repeats inside and across requests
        table = E.RecycleTable([[]], V)
        ver = E.Verifier(B, T)
        want = O.row_topk(logits, 8)
        for rep in range(2):
            out = ver.verify(logits, _dev_i32(tokens), None, n_nodes=_dev_i32(n_nodes), recycle=table)
            torch.cuda.synchronize()
            got = out["topk"].cpu().numpy()
            for b in range(B):
                assert np.array_equal(got[b, :n_nodes[b]], want[b, :n_nodes[b]]), (padded, b)
                assert (got[b, n_nodes[b]:] == -1).all()
        lf = logits.float()
        for b in range(min(B, 8)):
            nn = int(n_nodes[b])
            vals = torch.gather(lf[b, :nn], 1, torch.as_tensor(got[b, :nn], device="cuda", dtype=torch.long))
            ref = lf[b, :nn].topk(8).values
            assert torch.equal(torch.nan_to_num(vals, nan=1e30), torch.nan_to_num(ref, nan=1e30))
        cache = {}
        for b in range(B):
            O.recycle_update(cache, tokens[b, :n_nodes[b]], want[b, :n_nodes[b]])
        assert table.as_dict() == cache
        assert (table.owner == -1).all()
        ver.close()


# --------------------------------------------------------------------------------------
def test_edge_cases_and_error_behaviour():
    """Empty appends, a masked reset, an arena that fills up (flagged, never written past its end, and continued
    exactly after samd_dyn_grow), and the status-code / message behaviour of bad arguments."""
    E, K = _engine_mod()
    from samd_b200 import synth
    B = 6
    streams = [synth.copy_mix(400, 50, 900 + r) for r in range(B)]
    oracles = [O.Automaton() for _ in range(B)]
    dyn = E.DynSamBatch(B, 256)                                               # too small on purpose
    eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=8, len_bias=5, len_threshold=3)
    tok = np.stack([s[:200] for s in streams]).astype(np.int32)
    cnt = np.array([200, 0, 200, 13, 200, 1], dtype=np.int32)                  # ragged, two (nearly) empty
    eng.step(_dev_i32(tok), _dev_i32(cnt), None)
    for r in range(B):
        oracles[r].extend(streams[r][:cnt[r]])
    st = _dev_i32(np.array([int(s[200]) for s in streams]))
    eng.step(None, None, st)                                                  # lookup only
    torch.cuda.synchronize()
    for r in range(B):
        typ, seq, info = O.select_samd(oracles[r], None, int(st[r].item()), 8, 5, 3)
        assert eng.match_dyn[r].item() == info["match_dyn"]
    # masked reset: requests 0 and 3 start over, the others keep their automata
    mask = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.uint8, device="cuda")
    dyn.reset(mask)
    for r in (0, 3):
        oracles[r] = O.Automaton()
    # fill request 2 beyond its capacity: the overflow is flagged and nothing else is disturbed
    more = np.stack([s[200:300] for s in streams]).astype(np.int32)
    cnt2 = np.array([5, 5, 100, 5, 5, 5], dtype=np.int32)
    snap = E.DynSamBatch(B, 256)
    snap.copy_from(dyn)
    eng.step(_dev_i32(more), _dev_i32(cnt2), None)
    torch.cuda.synchronize()
    assert dyn.stats()["overflowed"] >= 1
    # the caller's recovery: grow the snapshot and replay the step - identical to an arena that was large enough
    big = snap.grown(1024)
    eng2 = E.DraftEngine(big, None, K.FLAVOUR_SAMD, n_predicts=8, len_bias=5, len_threshold=3)
    eng2.step(_dev_i32(more), _dev_i32(cnt2), None)
    for r in range(B):
        oracles[r].extend(streams[r][200:200 + cnt2[r]])
    torch.cuda.synchronize()
    assert big.stats()["overflowed"] == 0
    for r in range(B):
        ex = big.export(r, with_text=False)
        assert ex["n_states"] == oracles[r].n_states and ex["n_edges"] == oracles[r].n_edges, r
        assert np.array_equal(ex["link"], np.array(oracles[r].link))
    # bad arguments: a status code and a message, never a crash
    ver = E.Verifier(2, 4)
    lg = torch.zeros(3, 4, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(K.SamdError, match="batch exceeds"):
        ver.verify(lg, torch.zeros(3, 4, dtype=torch.int32, device="cuda"), None)
    lg = torch.zeros(2, 5, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(K.SamdError, match="n_nodes exceeds"):
        ver.verify(lg, torch.zeros(2, 5, dtype=torch.int32, device="cuda"), None)
    with pytest.raises(K.SamdError, match="dtype"):
        ver.verify(torch.zeros(2, 4, 64, dtype=torch.float64, device="cuda"), torch.zeros(2, 4, dtype=torch.int32, device="cuda"), None)
    with pytest.raises(K.SamdError, match="at least 8"):
        ver.verify(torch.zeros(2, 4, 5, dtype=torch.bfloat16, device="cuda"), torch.zeros(2, 4, dtype=torch.int32, device="cuda"),
                   None, want_topk=True)
    with pytest.raises(K.SamdError):
        E.DynSamBatch(0, 16)


def test_c3_static_and_dynamic_against_c_oracle():
    """BASELINE config 3's shape at a corpus the host builds in seconds (3 M tokens, ~5 M states): 512 persistent
    cursors, 6 steps of 1-8 tokens + lookup, the samd selection rule over the dynamic AND the static automaton
    (len_bias applied to the static match, ties to the dynamic one) - source, match lengths, state indices and
    16-token drafts against the oracle's C restatement, and the automaton's size against the oracle's."""
    E, K = _engine_mod()
    from c_oracle import CSam
    from samd_b200 import synth
    vocab, Q, steps = 32000, 512, 6
    docs = synth.make_corpus(3_000_000, vocab, 3100, singletons=True)
    st = E.StaticSamDevice.build(docs, synth.EOS, with_counts=False)
    ost = CSam.build(docs, synth.EOS)
    assert st.n_states == ost.info()["n_states"] and st.n_edges == ost.info()["n_edges"]
    q = synth.corpus_queries(docs, Q, 8 * steps + 1, vocab, 3101).astype(np.int32)
    rng = np.random.default_rng(3102)
    counts = rng.integers(1, 9, size=(steps, Q)).astype(np.int32)
    # expected, query by query (the C oracle's static automaton has one cursor)
    want = {}
    for r in range(Q):
        dyn = CSam(8 * steps + 16)
        ost.reset_cursor()
        pos = 0
        for s in range(steps):
            chunk = q[r, pos:pos + counts[s, r]]
            dyn.extend(chunk)
            ost.advance(chunk)
            pos += counts[s, r]
            want[(s, r)] = dyn.select_samd(ost, int(q[r, pos]), 16, 2, 4)
    dynb = E.DynSamBatch(Q, 8 * steps + 16)
    eng = E.DraftEngine(dynb, st, K.FLAVOUR_SAMD, n_predicts=16, len_bias=2, len_threshold=4)
    pos = np.zeros(Q, dtype=np.int64)
    seen = set()
    for s in range(steps):
        tok = np.zeros((Q, 8), dtype=np.int32)
        for r in range(Q):
            tok[r, :counts[s, r]] = q[r, pos[r]:pos[r] + counts[s, r]]
        pos += counts[s]
        start = q[np.arange(Q), pos]
        eng.step(_dev_i32(tok), _dev_i32(counts[s]), _dev_i32(start))
        torch.cuda.synchronize()
        typ, draft = eng.out_type.cpu().numpy(), eng.draft.cpu().numpy()
        md, ms = eng.match_dyn.cpu().numpy(), eng.match_static.cpu().numpy()
        idd, ids = eng.index_dyn.cpu().numpy(), eng.index_static.cpu().numpy()
        for r in range(Q):
            kind, seq, info = want[(s, r)]
            assert kind == typ[r], (s, r)
            assert (info[0], info[1]) == (idd[r], md[r]), (s, r)
            assert (info[2], info[3]) == (ids[r], ms[r]), (s, r)
            if kind != 2:
                assert seq == draft[r].tolist(), (s, r)
            seen.add(kind)
    assert {0, 1} <= seen                                      # both automata supplied drafts


def test_peer_exchange_single_rank_matches_key_path():
    """The NVLink peer-exchange kernels (csrc/xchg.cu) with one rank: the protocol (epoch parity, flags, re-armed key
    buffers, cursor advance fused into the look-up launch) must give what look-up keys + draft-from-keys give, step
    after step, also when the steps are captured in a CUDA graph.  (Several ranks: tools/p2p_check.py, bench.py.)"""
    E, K = _engine_mod()
    from samd_b200 import dist as D, synth
    docs = [d.tolist() for d in synth.make_corpus(40000, 500, 71, doc_len=(16, 96), singletons=True)]
    nq, qlen, steps = 96, 8 * 6 + 1, 6
    q = synth.corpus_queries([np.array(d) for d in docs], nq, qlen, 500, 72).astype(np.int32)
    a = D.ShardedStaticSam(docs, synth.EOS, 0, 1, nq, torch.device("cuda"))
    b = D.ShardedStaticSam(docs, synth.EOS, 0, 1, nq, torch.device("cuda"))
    assert b.connect_peers() and b.peers_ok()
    rng = np.random.default_rng(73)
    counts = rng.integers(0, 9, size=(steps, nq)).astype(np.int32)
    pos = np.zeros(nq, dtype=np.int64)
    toks, starts = [], []
    for s in range(steps):
        t = np.zeros((nq, 8), dtype=np.int32)
        for i in range(nq):
            t[i, :counts[s, i]] = q[i, pos[i]:pos[i] + counts[s, i]]
        pos += counts[s]
        toks.append(_dev_i32(t))
        starts.append(_dev_i32(q[np.arange(nq), pos]))
    d_cnt = [_dev_i32(c) for c in counts]
    want = [tuple(x.clone() for x in a.lookup_draft(starts[s], 16, tokens=toks[s], counts=d_cnt[s])) for s in range(steps)]
    out = (torch.empty(nq, dtype=torch.int32, device="cuda"), torch.empty(nq, 16, dtype=torch.int32, device="cuda"))
    for s in range(steps):                                       # eager
        m, d = b.lookup_draft(starts[s], 16, p2p=True, out=out, tokens=toks[s], counts=d_cnt[s])
        torch.cuda.synchronize()
        assert torch.equal(m, want[s][0]) and torch.equal(d, want[s][1]), s
    b.reset()
    outs = [(torch.empty_like(out[0]), torch.empty_like(out[1])) for _ in range(steps)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):                                     # all steps in one graph, replayed twice
        for s in range(steps):
            b.lookup_draft(starts[s], 16, p2p=True, out=outs[s], tokens=toks[s], counts=d_cnt[s])
    for rep in range(2):
        b.reset()
        g.replay()
        torch.cuda.synchronize()
        for s in range(steps):
            assert torch.equal(outs[s][0], want[s][0]) and torch.equal(outs[s][1], want[s][1]), (rep, s)
    assert b.peers_ok()


def test_static_l2_window_does_not_change_results():
    """samd_static_set_l2_window pins the record prefix in L2 through an access-policy window (a residency hint; measured
    on the 50 M-token automaton: no gain): lookups and drafts with and without it must be identical."""
    E, K = _engine_mod()
    from samd_b200 import synth
    docs = synth.make_corpus(300000, 32000, 71, singletons=True)
    st = E.StaticSamDevice.build(docs, synth.EOS)
    nq = 256
    q = synth.corpus_queries(docs, nq, 40, 32000, 72).astype(np.int32)

    def run():
        dyn = E.DynSamBatch(nq, 128)
        eng = E.DraftEngine(dyn, st, K.FLAVOUR_SAMD, n_predicts=16, len_bias=0, len_threshold=0)
        outs = []
        for s in range(4):
            eng.step(_dev_i32(q[:, 8 * s:8 * s + 8]), None, _dev_i32(q[:, 8 * s + 8]))
            outs.append(eng.out_buf.clone())
        torch.cuda.synchronize()
        return torch.stack(outs)

    base = run()
    st.set_l2_window(32 << 20)
    pinned = run()
    st.set_l2_window(0)
    again = run()
    assert torch.equal(base, pinned) and torch.equal(base, again)


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_words", [1, 3, 4, 5, 4096, 10240, 10243, 1 << 20])
def test_stage_copy_kernel(n_words):
    """samd_stage_copy (the host-buffer path's staging kernel): pinned host -> device and device -> pinned host, sizes
    with and without a 16-byte tail; unaligned pointers and sizes that are not a multiple of 4 are refused."""
    E, K = _engine_mod()
    src = torch.arange(n_words, dtype=torch.int32).mul_(2654435761 % 65521).pin_memory()
    dst = torch.zeros(n_words + 8, dtype=torch.int32, device="cuda")
    K.check(K.lib().samd_stage_copy(dst.data_ptr(), src.data_ptr(), n_words * 4, K.stream_ptr()))
    back = torch.zeros(n_words, dtype=torch.int32).pin_memory()
    K.check(K.lib().samd_stage_copy(back.data_ptr(), dst.data_ptr(), n_words * 4, K.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(dst[:n_words].cpu(), src) and int(dst[n_words:].abs().sum()) == 0
    assert torch.equal(back, src)
    assert K.lib().samd_stage_copy(dst.data_ptr() + 4, src.data_ptr(), 16, K.stream_ptr()) != 0
    assert K.lib().samd_stage_copy(dst.data_ptr(), src.data_ptr(), 6, K.stream_ptr()) != 0


# --------------------------------------------------------------------------------------
def test_counts_outside_the_token_row_are_clamped():
    """A per-request count below 0 or beyond the row of the token block is used as 0 / as the row length: the kernel never
    reads another request's tokens or past the block (both step kernel variants)."""
    E, K = _engine_mod()
    from samd_b200 import synth
    B, N = 6, 300
    streams = [synth.copy_mix(N + 64, 400, 8100 + r) for r in range(B)]
    for variant in (1, 0):
        K.lib().samd_step_set_variant(variant)
        try:
            engs = []
            for _ in range(2):
                dyn = E.DynSamBatch(B, N + 64)
                eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
                eng.step(_dev_i32(np.stack([s[:N] for s in streams])), None, None)
                engs.append(eng)
            tok = _dev_i32(np.stack([s[N:N + 8] for s in streams]))
            start = _dev_i32(np.array([s[N + 8] for s in streams]))
            wild = _dev_i32(np.array([100, -3, 8, 0, 1 << 30, -(1 << 31)], dtype=np.int64).astype(np.int32))
            tame = _dev_i32(np.array([8, 0, 8, 0, 8, 0]))
            engs[0].step(tok, wild, start)
            engs[1].step(tok, tame, start)
            torch.cuda.synchronize()
            assert torch.equal(engs[0].out_buf, engs[1].out_buf)
            assert np.array_equal(engs[0].dyn.meta()[:, :8], engs[1].dyn.meta()[:, :8])
        finally:
            K.lib().samd_step_set_variant(1)
