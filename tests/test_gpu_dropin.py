"""GPU tests of the drop-in `samd` / `samd_sam_only` packages: the reference's class surface
 driven the way the reference's callers drive it, checked against the golden
outputs of the reference's own classes."""
import os

import numpy as np
import pytest
import torch

import samd_oracle as O
from helpers import GOLDEN, load, unragged, docs_of

pytestmark = pytest.mark.gpu


def test_dyn_sam_class_api_and_growth():
    """samd.sam.DynSAM / samd_sam_only.sam.DynSAM method by method, starting from a tiny arena so
    that capacity growth (samd_dyn_grow) is exercised several times."""
    from samd.sam import DynSAM as DynA
    from samd_sam_only.sam import DynSAM as DynB
    z = load("dyn_sam.npz")
    for name in ("v3", "mix1k"):
        stream, cuts = z[f"{name}/stream"], z[f"{name}/cuts"]
        so = unragged(z[f"{name}/draft_so_flat"], z[f"{name}/draft_so_offs"])
        a, b = DynA(16), DynB(40, 4.0, "cuda")
        a._capacity = b._capacity = 64
        lo = 0
        for k, hi in enumerate(cuts):
            a.add_tokens(stream[lo:hi].tolist())
            b.add_tokens(stream[lo:hi].tolist())
            lo = hi
            tok = int(stream[hi])
            i, l = a.lookup(tok)
            assert (i, l) == b.lookup(tok) == (z[f"{name}/index"][k], z[f"{name}/match"][k])
            a.n_predicts = 16
            assert a.gen_draft(i, tok) == z[f"{name}/draft16"][k].tolist()
            a.n_predicts = 40
            assert a.gen_draft(i, tok) == z[f"{name}/draft40"][k].tolist()
            seq, buf = b.gen_draft(i, l, tok)
            assert seq == so[k] and buf["seq_position_ids"].tolist() == [list(range(len(seq)))]
        assert a._capacity > 64
        assert [a.cur_index, a.cur_length] == z[f"{name}/cursor"].tolist()
        assert a.max_length == cuts[-1] and a.input_ids[1:] == stream[:cuts[-1]].tolist()
        states = a.states
        assert [s.link for s in states] == z[f"{name}/link"].tolist()
        assert [s.min_endpos for s in states] == z[f"{name}/min_endpos"].tolist()
        ora = O.Automaton()
        ora.extend(stream[:cuts[-1]])
        assert [list(s.next.items()) for s in states] == [list(d.items()) for d in ora.trans]   # incl. insertion order
        a.reset()
        assert a.lookup(int(stream[0])) == (0, 0) and a.max_length == 0


def test_static_sam_class_api_and_pickles(tmp_path):
    """StaticSAM.build / reset / transfer_tokens / lookup / gen_draft, dump_sam + load_sam in the flat
    format, and load_sam of pickles written by the REFERENCE's dump_sam."""
    import samd.sam as A
    import samd_sam_only.sam as B
    z = load("static_sam.npz")
    name = "small"
    docs = docs_of(z, name)
    eos = int(z[f"{name}/eos"])
    built_a, built_b = A.build_sam(docs, eos), B.build_sam(docs, eos)
    A.dump_sam(str(tmp_path / "a.bin"), built_a)
    B.dump_sam(str(tmp_path / "b.bin"), built_b)
    variants_a = [built_a, A.load_sam(str(tmp_path / "a.bin")), A.load_sam(os.path.join(GOLDEN, "ref_static_samd.pkl"))]
    variants_b = [built_b, B.load_sam(str(tmp_path / "b.bin")), B.load_sam(os.path.join(GOLDEN, "ref_static_sam_only.pkl"))]
    q, steps = z[f"{name}/queries"], z[f"{name}/steps"]
    ora = O.build_static(docs, eos, count_occurrences=True)
    topk = O.build_topk(ora, 8)
    trees = unragged(z[f"{name}/tree_tok_flat"], z[f"{name}/tree_offs"])
    rets = unragged(z[f"{name}/tree_ret_flat"], z[f"{name}/tree_ret_offs"])
    for sa, sb in zip(variants_a, variants_b):
        sa.n_predicts = 16
        sb.max_predicts, sb.alpha, sb.K = 40, 4.0, 8
        prev_q, pos = -1, 0
        for k, (qi, p) in enumerate(steps[:120]):
            if qi != prev_q:
                sa.reset()
                sb.reset()
                prev_q, pos = qi, 0
            sa.transfer_tokens(q[qi, pos:p].tolist())
            sb.transfer_tokens(q[qi, pos:p].tolist())
            pos = p
            tok = int(q[qi, p])
            i, l = sa.lookup(tok)
            assert (i, l) == sb.lookup(tok) == (z[f"{name}/index"][k], z[f"{name}/match"][k])
            assert sa.gen_draft(i, tok) == z[f"{name}/draft16"][k].tolist()
            tree, buf = sb.gen_draft(i, max(l - 2, 0), tok)
            assert tree == trees[k]
            off = z[f"{name}/tree_offs"][k]
            assert buf["tree_position_ids"][0].tolist() == z[f"{name}/tree_depth_flat"][off:off + len(tree)].tolist()
            assert buf["tree_retrieve_indices"].reshape(-1).tolist() == rets[k]
            if k < 8:
                par = O.static_tree_sam_only(ora, topk, i, max(l - 2, 0), tok, 40, 4.0, 8)[1]
                assert np.array_equal(buf["tree_attn_mask"][0, 0].cpu().numpy(), O.tree_buffers(par)[0])
    states = built_a.states
    assert [s.link for s in states] == z[f"{name}/link"].tolist()
    assert [s.cnt_endpos for s in built_b.states] == z[f"{name}/cnt_endpos"].tolist()


def test_draft_model_both_packages():
    """DraftModel.reset / update / lookup of samd and samd_sam_only against the reference's outputs."""
    import samd
    import samd_sam_only as so
    z = load("draft_select.npz")
    docs = docs_of(z)
    da = samd.DraftModel(samd.SamdConfig(n_predicts=16, len_threshold=5, len_bias=5), sam_static=samd.build_sam(docs, 2),
                         device="cuda")
    db = so.DraftModel(so.SamdConfig(max_predicts=40, alpha=4.0, K=8, len_bias=5), sam_static=so.build_sam(docs, 2),
                       device="cuda")
    cuts = unragged(z["cuts_flat"], z["cuts_offs"])
    so_tok = unragged(z["so_tok_flat"], z["so_tok_offs"])
    so_ret = unragged(z["so_ret_flat"], z["so_ret_offs"])
    k = 0
    for r in range(2):
        stream = z["streams"][r]
        da.reset()
        db.reset()
        lo = 0
        for hi in cuts[r]:
            chunk = torch.as_tensor(stream[lo:hi]).cuda()
            da.update(tokens=chunk)
            db.update(tokens=chunk)
            lo = hi
            tok = int(stream[hi])
            kind, seq, _ = da.lookup(tok)
            if z["samd_source"][k] == 2:
                assert kind.value == "tree" and len(seq) == 61 and seq[0] == tok
            else:
                assert kind.value == "sequence" and seq == z["samd_seq"][k].tolist()
            kind2, toks2, buf2 = db.lookup(tok)
            assert toks2 == so_tok[k]
            if z["so_type"][k] == 0:
                assert kind2.value == "sequence"
            else:
                assert kind2.value == "tree" and buf2["tree_retrieve_indices"].reshape(-1).tolist() == so_ret[k]
            k += 1


@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32"])
def test_eval_posterior_gathered_logits(dt):
    """samd.utils.eval_posterior with the reference's calling convention ([P, D, V] gathered logits)."""
    from samd.utils import eval_posterior, SamdGenerationConfig
    z = load("verify.npz")
    bits = torch.from_numpy(z["bf16/logits_bits"]).view(torch.bfloat16)
    lg = {"bf16": bits, "fp16": bits.float().to(torch.float16), "fp32": bits.float()}[dt].cuda()
    key = "bf16" if dt == "fp32" else dt                     # fp32 values are exactly the bf16 ones
    ri = torch.as_tensor(z["retrieve"]).cuda()
    cfg = SamdGenerationConfig()
    for b in range(0, lg.shape[0], 3):
        ext = torch.cat([torch.as_tensor(z["tree_tokens"][b]), torch.zeros(1, dtype=torch.long)]).cuda()
        best, alen, sample_p = eval_posterior(lg[b][ri], ext[ri], cfg)
        assert int(best) == z[f"{key}/best"][b] and int(alen) == z[f"{key}/accept_len"][b]
        assert int(torch.argmax(sample_p, dim=-1)) == z[f"{key}/next_token"][b]


def test_static_cache_select_indices():
    """SamdStaticCache.update / select_indices / set_length / reset against the reference's cache."""
    from types import SimpleNamespace
    from samd.cache import SamdStaticCache
    z = load("verify.npz")
    init = torch.from_numpy(z["kv/init_bits"]).view(torch.bfloat16)          # [2L, 1, H, ML, DH]
    L2, _, H, ML, DH = init.shape
    cfg = SimpleNamespace(num_hidden_layers=L2 // 2, max_position_embeddings=ML, hidden_size=H * DH, num_attention_heads=H,
                          num_key_value_heads=H, head_dim=DH)
    cache = SamdStaticCache(cfg, batch_size=1, max_cache_len=ML, device="cuda", dtype=torch.bfloat16, hf_device_map={})
    for c, (b, start) in enumerate(z["kv/cases"][:8]):
        cache.reset()
        for l in range(L2 // 2):                                             # "prefill" writes rows [0, ML)
            k_view, _ = cache.update(init[l].cuda(), init[L2 // 2 + l].cuda(), l)
            assert k_view.shape[2] == ML
        cache.cache_length = int(start)
        al = int(z["bf16/accept_len"][b])
        cache.select_indices(torch.as_tensor(z["bf16/indices"][b][:al]).cuda(), al)
        assert cache.get_seq_length() == start + al
        got = torch.stack(cache.key_cache + cache.value_cache).view(torch.int16).cpu().numpy()
        assert np.array_equal(got, z["kv/after_bits"][c])
        cache.select_indices(None, 3)
        assert cache.cache_length == start + al + 3


def _tiny_llama(dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=96, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=2048, attn_implementation="eager")
    return LlamaForCausalLM(cfg).to("cuda", dtype).eval()


@torch.inference_mode()
def _plain_greedy(lm, ids, n_new):
    out = ids.clone()
    for _ in range(n_new):
        nxt = lm(input_ids=out).logits[:, -1].argmax(-1, keepdim=True)
        out = torch.cat([out, nxt], dim=1)
    return out[0].tolist()


@pytest.mark.parametrize("pkg", ["samd_sam_only", "samd"])
def test_generate_is_lossless_on_a_tiny_llama(pkg):
    """The reference's own acceptance criterion (evaluation/equal.py): speculative greedy decoding
    must reproduce plain greedy decoding token for token.  fp32 tiny random Llama, repetitive prompt
    (so that drafts get accepted), both packages; the sam_only run uses a static SAM so that tree
    drafts, the 4-D tree mask and the fused KV compaction are exercised."""
    import importlib
    from samd_b200 import synth
    mod = importlib.import_module(pkg)
    lm = _tiny_llama(torch.float32)
    prompt = torch.as_tensor(synth.copy_mix(160, 96, 77, p_copy=0.7)[None]).cuda()
    n_new = 96
    want = _plain_greedy(lm, prompt, n_new)
    docs = [want[40:120], want[100:200], synth.copy_mix(300, 96, 5).tolist()] + [[i] for i in range(96)]
    if pkg == "samd":
        cfg = mod.SamdConfig(n_predicts=8, len_threshold=3, len_bias=1)
    else:
        cfg = mod.SamdConfig(max_predicts=16, alpha=4.0, K=4, len_bias=0)
    draft = mod.DraftModel(cfg, sam_static=mod.build_sam(docs, 2), lm=lm, dtype=torch.float32, device="cuda")
    model = mod.SamdModel(cfg, lm, draft, eos_token_id=-1, dtype=torch.float32, device="cuda")
    out = model.generate(prompt, generation_config=mod.SamdGenerationConfig(max_new_tokens=n_new, max_cache_len=512))
    assert out.output_ids[0] == want[:len(out.output_ids[0])]
    assert len(out.output_ids[0]) == prompt.shape[1] + n_new
    assert sum(out.accepet_length_per_step) >= n_new and out.decode_steps < n_new      # some drafts were accepted
    assert out.decode_tokens == sum(out.accepet_length_per_step)


def test_generate_with_sampling_on_a_tiny_llama():
    """SamdGenerationConfig(greedy=False) - the reference's typical-acceptance branch (samd/utils.py:142-184) through
    samd_verify_sample, end to end: the loop runs to the requested length with consistent bookkeeping and is a pure
    function of (torch seed, Philox seed / offset).  (Like the reference, only the ACCEPTANCE test uses the temperature /
    top-p / top-k processor; the first token and the non-residual next tokens are drawn from the plain softmax -
    samd/samd_model.py:127, samd/utils.py:176-179 - so a low temperature does not reproduce greedy decoding.)"""
    import samd as mod
    from samd import utils as U
    from samd_b200 import synth
    lm = _tiny_llama(torch.float32)
    prompt = torch.as_tensor(synth.copy_mix(160, 96, 77, p_copy=0.7)[None]).cuda()
    n_new = 64
    cfg = mod.SamdConfig(n_predicts=8, len_threshold=3, len_bias=1)
    runs = []
    for _ in range(2):
        # (a fresh draft model per run: like the reference's, the Token-Recycle table survives DraftModel.reset())
        draft = mod.DraftModel(cfg, lm=lm, dtype=torch.float32, device="cuda")
        model = mod.SamdModel(cfg, lm, draft, eos_token_id=-1, dtype=torch.float32, device="cuda")
        torch.manual_seed(0)
        U._sampling.clear()                                              # a fresh Philox stream (offset 0)
        out = model.generate(prompt, generation_config=mod.SamdGenerationConfig(max_new_tokens=n_new, max_cache_len=512,
                                                                                greedy=False, temperature=0.8, top_p=0.95,
                                                                                top_k=20, seed=4))
        ids = out.output_ids[0]
        assert len(ids) == prompt.shape[1] + n_new and ids[:prompt.shape[1]] == prompt[0].tolist()
        assert out.decode_tokens == sum(out.accepet_length_per_step) and all(a >= 1 for a in out.accepet_length_per_step)
        assert all(0 <= t < 96 for t in ids)
        runs.append(ids)
    assert runs[0] == runs[1]


def test_batched_decoder_is_lossless_on_a_tiny_llama():
    """B = 4 requests with ragged prompts decoded in lockstep (one samd_step launch,
    one LM forward, one samd_verify_compact launch per step, per-request KV offsets) must reproduce plain
    greedy decoding of every request, including EOS truncation."""
    from samd_b200 import synth
    from samd_b200.batched import BatchedSamdDecoder
    lm = _tiny_llama(torch.float32)
    prompts = [synth.copy_mix(n, 96, 300 + i, p_copy=0.7).tolist() for i, n in enumerate((150, 97, 200, 64))]
    n_new = 64
    want = [_plain_greedy(lm, torch.as_tensor([p]).cuda(), n_new)[len(p):] for p in prompts]
    dec = BatchedSamdDecoder(lm, 4, 512, n_predicts=8, len_bias=5, len_threshold=3, dtype=torch.float32)
    got, stats = dec.generate(prompts, n_new)
    assert got == want
    assert stats["steps"] < n_new                                  # drafts were accepted
    # the decode step is captured as a CUDA graph per key-length bucket and the host reads one flag every 4 steps
    assert stats["graphed"] and stats["graphs"] >= 1
    assert stats["host_syncs"] <= stats["steps_run"] // 4 + 1 and stats["steps_run"] - stats["steps"] < 4
    got_e, stats_e = dec.generate(prompts, n_new, graph=False)      # the same body, eagerly: same stream of tokens
    assert got_e == want and stats_e["accept_lengths"] == stats["accept_lengths"] and not stats_e["graphed"]
    eos = want[1][10]                                              # pick a token request 1 really emits
    dec2 = BatchedSamdDecoder(lm, 4, 512, n_predicts=8, len_bias=5, len_threshold=3, eos_token_id=eos, dtype=torch.float32)
    got2, _ = dec2.generate(prompts, n_new)
    for g, w in zip(got2, want):
        cut = w.index(eos) + 1 if eos in w else len(w)
        assert g == w[:cut]


def test_batched_decoder_with_token_recycle_tree_is_lossless():
    """SURVEY section 8f rows 2 + 3 together: requests whose suffix match is short draft the reference's 61-node
    Token-Recycle tree from the device table the verify launch maintains, the others draft sequences; per-request
    node counts, masks, positions and path tables; samd_verify_compact moves the accepted rows.  The output must be
    plain greedy decoding, and the tree must actually get accepted tokens."""
    from samd_b200 import synth, _cabi as K
    from samd_b200.batched import BatchedSamdDecoder
    lm = _tiny_llama(torch.float32)
    prompts = [synth.copy_mix(n, 96, 500 + i, p_copy=0.5).tolist() for i, n in enumerate((120, 90, 160, 70))]
    n_new = 48
    want = [_plain_greedy(lm, torch.as_tensor([p]).cuda(), n_new)[len(p):] for p in prompts]
    # a high threshold sends most steps to the tree
    dec = BatchedSamdDecoder(lm, 4, 512, n_predicts=8, len_bias=5, len_threshold=6, dtype=torch.float32,
                             tree=synth.token_recycle_tree())
    got, stats = dec.generate(prompts, n_new)
    assert got == want
    assert stats["steps"] < n_new
    assert len(dec.table.as_dict()) > 0
