"""Generate tests/golden/*.npz by running the REFERENCE's own classes (imported from
/root/reference through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container only:   python oracle/gen_golden.py
The fixtures are committed; the GPU box never needs the reference checkout.
Each fixture stores its inputs next to the reference's outputs, so a test can feed the
same inputs to the oracle restatement and to the CUDA path.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "sam-decoding_b200"))

import ref_loader  # noqa: E402
from samd_b200 import synth  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def ragged(rows):
    flat = np.array([x for r in rows for x in r], dtype=np.int64)
    offs = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    return flat, offs


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# --------------------------------------------------------------------------------------
def dyn_fixture(ns):
    """DynSAM (both flavours): chunked add_tokens, then lookup + gen_draft with the next token."""
    cases = []
    specs = [("v2", 2, 96, 11), ("v3", 3, 128, 12), ("v4", 4, 160, 13), ("v6", 6, 200, 14),
             ("mix1k", 32000, 1024, 1000), ("mix4k", 32000, 4096, 1001), ("light2k", 32000, 2048, 1002)]
    out = {}
    names = []
    for name, vocab, n, seed in specs:
        if vocab <= 8:
            stream = synth.small_alphabet(n + 1, vocab, seed)
        else:
            stream = synth.copy_mix(n + 1, vocab, seed, uniform_fresh=name.startswith("light"))
        rng = np.random.default_rng(seed + 99)
        prompt = n // 2 if vocab > 8 else 1
        cuts = [prompt]
        while cuts[-1] < n:
            cuts.append(min(n, cuts[-1] + int(rng.integers(1, 9))))
        a = ns.samd_sam.DynSAM(16)
        b = ns.so_sam.DynSAM(40, 4.0, "cpu")
        idx, mlen, d16, d40, dso = [], [], [], [], []
        lo = 0
        for hi in cuts:
            chunk = stream[lo:hi].tolist()
            a.add_tokens(chunk)
            b.add_tokens(chunk)
            lo = hi
            tok = int(stream[hi])
            i, l = a.lookup(tok)
            i2, l2 = b.lookup(tok)
            assert (i, l) == (i2, l2)
            idx.append(i)
            mlen.append(l)
            a.n_predicts = 16
            d16.append(a.gen_draft(i, tok))
            a.n_predicts = 40
            d40.append(a.gen_draft(i, tok))
            dso.append(b.gen_draft(i2, l2, tok)[0])
        out[f"{name}/stream"] = stream
        out[f"{name}/cuts"] = np.array(cuts, dtype=np.int64)
        out[f"{name}/index"] = np.array(idx, dtype=np.int64)
        out[f"{name}/match"] = np.array(mlen, dtype=np.int64)
        out[f"{name}/draft16"] = np.array(d16, dtype=np.int64)
        out[f"{name}/draft40"] = np.array(d40, dtype=np.int64)
        f, o = ragged(dso)
        out[f"{name}/draft_so_flat"], out[f"{name}/draft_so_offs"] = f, o
        out[f"{name}/link"] = np.array([s.link for s in a.states], dtype=np.int64)
        out[f"{name}/length"] = np.array([s.length for s in a.states], dtype=np.int64)
        out[f"{name}/min_endpos"] = np.array([s.min_endpos for s in a.states], dtype=np.int64)
        out[f"{name}/cursor"] = np.array([a.cur_index, a.cur_length], dtype=np.int64)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "dyn_sam.npz"), **out)
    print("dyn_sam.npz", len(out))


# --------------------------------------------------------------------------------------
def static_fixture(ns):
    """StaticSAM, both flavours, over small corpora incl. the vocab-singleton convention."""
    out = {}
    names = []
    specs = [("tiny", 5, 300, 21, False), ("small", 64, 3000, 22, True), ("mid", 2000, 20000, 23, True)]
    for name, vocab, n_tok, seed, singles in specs:
        if vocab <= 8:
            rng = np.random.default_rng(seed)
            docs = []
            left = n_tok
            while left > 0:
                k = int(rng.integers(3, 40))
                d = rng.integers(0, vocab, size=k).astype(np.int64)
                docs.append(d)
                left -= k
            eos = 2
        else:
            docs = synth.make_corpus(n_tok, vocab, seed, doc_len=(16, 96), singletons=singles)
            eos = synth.EOS
        doc_lists = [d.tolist() for d in docs]
        sa = ns.samd_sam.StaticSAM.build(doc_lists, eos, verbose=False)
        with quiet():
            sb = ns.so_sam.StaticSAM.build(doc_lists, eos, verbose=False)
        sb.device = "cpu"
        f, o = ragged(doc_lists)
        out[f"{name}/docs_flat"], out[f"{name}/docs_offs"] = f, o
        out[f"{name}/eos"] = np.array(eos)
        out[f"{name}/vocab"] = np.array(vocab)
        out[f"{name}/link"] = np.array([s.link for s in sa.states], dtype=np.int64)
        out[f"{name}/length"] = np.array([s.length for s in sa.states], dtype=np.int64)
        out[f"{name}/min_endpos"] = np.array([s.min_endpos for s in sa.states], dtype=np.int64)
        out[f"{name}/cnt_endpos"] = np.array([s.cnt_endpos for s in sb.states], dtype=np.int64)
        out[f"{name}/n_edges"] = np.array(sum(len(s.next) for s in sa.states))
        tk_tok = np.full((len(sb.states), 8), -1, dtype=np.int64)
        tk_idx = np.full((len(sb.states), 8), -1, dtype=np.int64)
        for i, lst in enumerate(sb.states_topk_next):
            for j, (t, s) in enumerate(lst):
                tk_tok[i, j], tk_idx[i, j] = t, s
        out[f"{name}/topk_tok"], out[f"{name}/topk_idx"] = tk_tok, tk_idx
        # queries: windows walked with transfer_tokens in chunks, then lookup + gen_draft
        nq = 24
        qlen = 48
        if vocab <= 8:
            q = np.random.default_rng(seed + 5).integers(0, vocab, size=(nq, qlen)).astype(np.int64)
        else:
            q = synth.corpus_queries(docs, nq, qlen, vocab, seed + 5)
        out[f"{name}/queries"] = q
        steps = []
        r_idx, r_len, r_d16, tree_tok, tree_par_flat = [], [], [], [], []
        tree_ret = []
        rng = np.random.default_rng(seed + 6)
        for qi in range(nq):
            sa.reset()
            sb.reset()
            pos = 0
            while pos < qlen - 1:
                k = int(rng.integers(1, 9))
                k = min(k, qlen - 1 - pos)
                chunk = q[qi, pos:pos + k].tolist()
                sa.transfer_tokens(chunk)
                sb.transfer_tokens(chunk)
                pos += k
                tok = int(q[qi, pos])
                i, l = sa.lookup(tok)
                assert (i, l) == sb.lookup(tok)
                sa.n_predicts = 16
                steps.append((qi, pos))
                r_idx.append(i)
                r_len.append(l)
                r_d16.append(sa.gen_draft(i, tok))
                sb.max_predicts, sb.alpha, sb.K = 40, 4.0, 8
                toks, buf = sb.gen_draft(i, max(l - 2, 0), tok)     # bias 2 keeps trees non-trivial
                tree_tok.append(toks)
                tree_par_flat.append(buf["tree_position_ids"][0].tolist())
                tree_ret.append(buf["tree_retrieve_indices"].numpy())
        out[f"{name}/steps"] = np.array(steps, dtype=np.int64)
        out[f"{name}/index"] = np.array(r_idx, dtype=np.int64)
        out[f"{name}/match"] = np.array(r_len, dtype=np.int64)
        out[f"{name}/draft16"] = np.array(r_d16, dtype=np.int64)
        f, o = ragged(tree_tok)
        out[f"{name}/tree_tok_flat"], out[f"{name}/tree_offs"] = f, o
        out[f"{name}/tree_depth_flat"] = ragged(tree_par_flat)[0]
        rf, ro = ragged([r.reshape(-1).tolist() for r in tree_ret])
        out[f"{name}/tree_ret_flat"], out[f"{name}/tree_ret_offs"] = rf, ro
        out[f"{name}/tree_ret_shape"] = np.array([r.shape for r in tree_ret], dtype=np.int64)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "static_sam.npz"), **out)
    print("static_sam.npz", len(out))


# --------------------------------------------------------------------------------------
def select_fixture(ns):
    """DraftModel.lookup / update (both packages) driven like SamdModel does: prefill
    update, then per step lookup(next token) and update(accepted chunk)."""
    out = {}
    vocab = 2000
    docs = synth.make_corpus(20000, vocab, 23, doc_len=(16, 96), singletons=True)
    doc_lists = [d.tolist() for d in docs]
    f, o = ragged(doc_lists)
    out["docs_flat"], out["docs_offs"] = f, o
    out["vocab"] = np.array(vocab)
    flat = np.concatenate([d for d in docs if len(d) > 1])
    with quiet():
        cfg_a = ns.samd_config.SamdConfig(n_predicts=16, len_threshold=5, len_bias=5)
        sa = ns.samd_sam.StaticSAM.build(doc_lists, synth.EOS, verbose=False)
        sb = ns.so_sam.StaticSAM.build(doc_lists, synth.EOS, verbose=False)
        cfg_b = ns.so_config.SamdConfig(max_predicts=40, alpha=4.0, K=8, len_bias=5)
        da = ns.samd_draft.DraftModel(cfg_a, sam_static=sa, lm=None, device="cpu")
        db = ns.so_draft.DraftModel(cfg_b, sam_static=sb, lm=None, device="cpu")
    n_req = 6
    streams, cuts_all = [], []
    typ_a, seq_a, src_a = [], [], []
    typ_b, tok_b, ret_b = [], [], []
    for r in range(n_req):
        stream = synth.copy_mix(700, vocab, 3000 + r, source=flat, p_source=0.5)
        rng = np.random.default_rng(4000 + r)
        cuts = [256]
        while cuts[-1] < 690:
            cuts.append(min(690, cuts[-1] + int(rng.integers(1, 9))))
        da.reset()
        db.reset()
        lo = 0
        for hi in cuts:
            chunk = torch.tensor(stream[lo:hi])
            da.sam_dyn.add_tokens(chunk.tolist())          # DraftModel.update minus the tree model
            da.sam_static.transfer_tokens(chunk.tolist())
            db.update(tokens=chunk)
            lo = hi
            tok = int(stream[hi])
            # samd flavour (tree fallback is out of scope: record only the decision)
            i_d, m_d = da.sam_dyn.lookup(tok)
            i_s, m_s = da.sam_static.lookup(tok)
            t, seq, _ = da.lookup(tok)
            if t.value == "sequence":
                typ_a.append(0)
                seq_a.append(seq)
                src_a.append(0 if m_d >= m_s - 5 else 1)
            else:
                typ_a.append(1)
                seq_a.append([0] * 16)
                src_a.append(2)
            t2, toks2, buf2 = db.lookup(tok)
            typ_b.append(0 if t2.value == "sequence" else 1)
            tok_b.append(toks2)
            ret_b.append(buf2["tree_retrieve_indices"].numpy().reshape(-1).tolist()
                         if t2.value == "tree" else [])
        streams.append(stream)
        cuts_all.append(cuts)
    out["streams"] = np.array(streams)
    cf, co = ragged(cuts_all)
    out["cuts_flat"], out["cuts_offs"] = cf, co
    out["samd_type"] = np.array(typ_a)
    out["samd_seq"] = np.array(seq_a, dtype=np.int64)
    out["samd_source"] = np.array(src_a)
    out["so_type"] = np.array(typ_b)
    out["so_tok_flat"], out["so_tok_offs"] = ragged(tok_b)
    out["so_ret_flat"], out["so_ret_offs"] = ragged(ret_b)
    np.savez_compressed(os.path.join(OUT, "draft_select.npz"), **out)
    print("draft_select.npz", len(out), "samd types", np.bincount(out["samd_source"]), "so types",
          np.bincount(out["so_type"]))


# --------------------------------------------------------------------------------------
def verify_fixture(ns):
    """Gather + eval_posterior (greedy) + update_state slices + select_indices on small shapes,
    including exact ties, NaNs, -0.0/+0.0 and padded (-1) retrieve entries."""
    out = {}
    tree = synth.token_recycle_tree()
    with quiet():
        ri_ref = ns.tr_utils.gen_buffers(tree, "cpu")["tree_retrieve_indices"]
    ri = synth.tree_retrieve_indices(tree)
    assert np.array_equal(ri_ref.numpy(), ri)
    out["retrieve"] = ri.astype(np.int64)
    gcfg = SimpleNamespace(greedy=True)
    vocab, T, B = 256, 61, 32
    rng = np.random.default_rng(77)
    tree_tokens = rng.integers(1, vocab, size=(B, T)).astype(np.int64)
    base = None
    for dt_name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        logits, _ = synth.planted_logits(B, T, vocab, tree_tokens, ri, seed=78, dtype="float32")
        # adversarial rows: exact ties, NaN, +-0, +inf, token-0 maxima (pad acceptance)
        for b in range(0, B, 6):
            logits[b, rng.integers(0, T), :] = 0.0
            logits[b, rng.integers(0, T), rng.integers(0, vocab)] = float("nan")
            r = int(rng.integers(0, T))
            logits[b, r, :] = -0.0
            logits[b, r, 7] = 0.0
        for b in range(3, B, 6):
            r = int(rng.integers(0, T))
            logits[b, r, 100] = float("inf")
            logits[b, r, 50] = float("inf")
            logits[b, :, 0] = 30.0 if b % 12 == 3 else logits[b, :, 0]
        if base is None:
            base = logits.to(torch.bfloat16)
        # fp16 logits are DERIVED from the stored bf16 bits (exact widening, IEEE narrowing) so
        # that only one copy of the bits is committed
        lg = base if dt is torch.bfloat16 else base.float().to(torch.float16)
        best, alen, nxt, toks, idxs = [], [], [], [], []
        for b in range(B):
            tokens_ext = torch.tensor(tree_tokens[b].tolist() + [0], dtype=torch.long)
            cand = tokens_ext[ri_ref]
            cand_logits = lg[b][ri_ref]                     # samd/samd_model.py:164
            bc, al, sp = ns.samd_utils.eval_posterior(cand_logits, cand, gcfg)
            al = int(al)
            best.append(int(bc))
            alen.append(al)
            nxt.append(int(torch.argmax(sp, dim=-1)))
            toks.append(cand[int(bc)][:al].tolist() + [-9] * (6 - al))
            idxs.append(ri_ref[int(bc)][:al].tolist() + [-9] * (6 - al))
        if dt is torch.bfloat16:
            out["bf16/logits_bits"] = lg.view(torch.int16).numpy()
        out[f"{dt_name}/best"] = np.array(best)
        out[f"{dt_name}/accept_len"] = np.array(alen)
        out[f"{dt_name}/next_token"] = np.array(nxt)
        out[f"{dt_name}/tokens"] = np.array(toks)
        out[f"{dt_name}/indices"] = np.array(idxs)
        out[f"{dt_name}/node_argmax"] = torch.argmax(lg, dim=-1).numpy()
    out["tree_tokens"] = tree_tokens
    # sequence-type candidates (P = 1, identity retrieve)
    n = 16
    seq_tok = rng.integers(1, vocab, size=(B, n)).astype(np.int64)
    ident = np.arange(n, dtype=np.int64)[None]
    lg, _ = synth.planted_logits(B, n, vocab, seq_tok, ident, seed=79, dtype="bfloat16")
    sb, sa, sn = [], [], []
    for b in range(B):
        cand = torch.tensor(seq_tok[b:b + 1])
        bc, al, sp = ns.samd_utils.eval_posterior(lg[b:b + 1], cand, gcfg)
        sb.append(int(bc))
        sa.append(int(al))
        sn.append(int(torch.argmax(sp, dim=-1)))
    out["seq/tokens"] = seq_tok
    out["seq/logits_bits"] = lg.view(torch.int16).numpy()
    out["seq/accept_len"] = np.array(sa)
    out["seq/next_token"] = np.array(sn)
    # KV compaction through the reference's own select_indices (unbound on a shim)
    L, H, ML, DH = 3, 2, 128, 8
    kv0 = torch.arange(2 * L * H * ML * DH, dtype=torch.float32).reshape(2 * L, 1, H, ML, DH)
    kv0 = (kv0 % 251).to(torch.bfloat16)
    cases = []
    after = []
    for c in range(16):
        b = int(rng.integers(0, B))
        start = int(rng.integers(1, 60))
        al = int(out["bf16/accept_len"][b])
        ind = out["bf16/indices"][b][:al]
        shim = SimpleNamespace(key_cache=[kv0[i].clone() for i in range(L)],
                               value_cache=[kv0[L + i].clone() for i in range(L)], cache_length=start,
                               last_length=start + T)
        ns.samd_cache.SamdStaticCache.select_indices(shim, torch.tensor(ind, dtype=torch.long), al)
        assert shim.cache_length == start + al
        cases.append((b, start))
        after.append(torch.stack(shim.key_cache + shim.value_cache).view(torch.int16).numpy())
    out["kv/init_bits"] = kv0.view(torch.int16).numpy()
    out["kv/cases"] = np.array(cases)
    out["kv/after_bits"] = np.array(after)
    np.savez_compressed(os.path.join(OUT, "verify.npz"), **out)
    print("verify.npz", len(out), "accept hist", np.bincount(out["bf16/accept_len"]))


# --------------------------------------------------------------------------------------
def loop_fixture(ns):
    """The decode loop of samd_sam_only/samd_model.py (prefill update -> {gen_candidates ->
    fake LM logits -> eval_posterior -> update_state}*), driven with the reference's own
    gen_candidates / eval_posterior / DraftModel and a fake LM that knows the continuation."""
    out = {}
    vocab = 96
    gcfg = ns.so_utils.SamdGenerationConfig(max_new_tokens=256, max_cache_len=4096)
    with quiet():
        cfg = ns.so_config.SamdConfig(max_predicts=40, alpha=4.0, K=8, len_bias=5)
    names = []
    for name, seed, plen in (("a", 501, 512), ("b", 502, 1024), ("c", 503, 64)):
        full = synth.copy_mix(plen + 700, vocab, seed, p_copy=0.6)
        prompt = full[:plen]
        with quiet():
            st = ns.so_sam.StaticSAM.build([[0]], 2, verbose=False)   # effectively empty static SAM
            st.device = "cpu"
            dm = ns.so_draft.DraftModel(cfg, sam_static=st, lm=None, device="cpu")
        dm.reset()
        dm.update(tokens=torch.tensor(prompt))
        pos = plen
        sample_p = torch.zeros(1, vocab)
        sample_p[0, int(full[pos])] = 1.0
        new_tokens, accepts = [], []
        while len(new_tokens) < gcfg.max_new_tokens:
            cands = ns.so_utils.gen_candidates(sample_p, None, dm, cfg, gcfg, "cpu")
            assert cands.type.value == "sequence"
            toks = cands.tokens[0].tolist()
            n = len(toks)
            logits = torch.zeros(1, n, vocab)
            for j in range(n):
                logits[0, j, int(full[pos + j + 1])] = 1.0
            bc, al, sample_p = ns.so_utils.eval_posterior(logits, cands.candidate_tokens, gcfg)
            acc = cands.candidate_tokens[bc][:al]
            dm.update(tokens=acc)
            acc = acc.tolist()
            new_tokens.extend(acc)
            accepts.append(len(acc))
            pos += len(acc)
        out[f"{name}/full"] = full
        out[f"{name}/plen"] = np.array(plen)
        out[f"{name}/new_tokens"] = np.array(new_tokens[:gcfg.max_new_tokens])
        out[f"{name}/accepts"] = np.array(accepts)
        names.append(name)
    out["names"] = np.array(names)
    out["vocab"] = np.array(vocab)
    np.savez_compressed(os.path.join(OUT, "decode_loop.npz"), **out)
    print("decode_loop.npz", {n: (len(out[f'{n}/accepts']), float(out[f'{n}/accepts'].mean())) for n in names})


def recycle_fixture(ns):
    """TokenRecycle.update / gen_draft (samd/tree_model/token_recycle/token_recycle.py) driven for a few decode
    steps: logits rows with pairwise-distinct values (so torch.topk's unspecified tie order cannot matter),
    tokens that repeat inside a step (last row wins) and across steps (later step wins), starts that miss."""
    import importlib
    tr = importlib.import_module("samd.tree_model.token_recycle.token_recycle")
    cfg = ns.samd_config.SamdConfig()
    model = tr.TokenRecycle(cfg, None, torch.float32, "cpu")
    tree = cfg.tree
    T, V, steps = len(tree), 640, 6
    rng = np.random.default_rng(77)
    # distinct finite bf16 bit patterns: positive 0x3C00..0x4600 and their negatives
    pool = np.concatenate([np.arange(0x3C00, 0x4600), np.arange(0xBC00, 0xC600)]).astype(np.uint16)
    out = {"tree_flat": ragged(tree)[0], "tree_offs": ragged(tree)[1], "vocab": np.array(V)}
    all_bits, all_tok, all_topk, drafts, starts = [], [], [], [], []
    for s in range(steps):
        bits = np.stack([rng.choice(pool, size=V, replace=False) for _ in range(T)])
        logits = torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16)
        tokens = rng.integers(3, 60, size=T)                       # small range: repeats inside and across steps
        model.update(tree_tokens=torch.as_tensor(tokens), tree_logits=logits.float())
        all_bits.append(bits)
        all_tok.append(tokens)
        all_topk.append(np.array(model.logits_to_topk(logits.float())))
        for q in range(4):
            st = int(rng.integers(3, 70))                          # some starts have no entry
            starts.append(st)
            drafts.append(model.gen_draft(st)[0])
    out["logits_bits"] = np.stack(all_bits)                         # [steps, T, V] bf16 bits
    out["tokens"] = np.stack(all_tok)
    out["topk"] = np.stack(all_topk)                                # [steps, T, 8]
    out["starts"] = np.array(starts).reshape(steps, 4)
    out["drafts"] = np.array(drafts).reshape(steps, 4, T)
    keys = sorted(model.cache)
    out["cache_keys"] = np.array(keys)
    out["cache_vals"] = np.array([model.cache[k] for k in keys])
    np.savez_compressed(os.path.join(OUT, "recycle.npz"), **out)
    print("recycle.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


def pickle_fixture(ns):
    """Pickles written by the reference's own dump_sam (samd/sam/utils.py:20-22), to check that this
    framework's load_sam reads them (module path / class names / attribute layout)."""
    z = np.load(os.path.join(OUT, "static_sam.npz"))
    flat, offs = z["small/docs_flat"], z["small/docs_offs"]
    docs = [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    sa = ns.samd_sam.StaticSAM.build(docs, int(z["small/eos"]), verbose=False)
    ns.samd_sam.dump_sam(os.path.join(OUT, "ref_static_samd.pkl"), sa)
    with quiet():
        sb = ns.so_sam.StaticSAM.build(docs, int(z["small/eos"]), verbose=False)
    ns.so_sam.dump_sam(os.path.join(OUT, "ref_static_sam_only.pkl"), sb)
    print("pickles", os.path.getsize(os.path.join(OUT, "ref_static_samd.pkl")),
          os.path.getsize(os.path.join(OUT, "ref_static_sam_only.pkl")))


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "pickles":
        pickle_fixture(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "recycle":
        recycle_fixture(ns)
        return
    dyn_fixture(ns)
    static_fixture(ns)
    select_fixture(ns)
    verify_fixture(ns)
    loop_fixture(ns)
    recycle_fixture(ns)
    pickle_fixture(ns)


if __name__ == "__main__":
    main()
