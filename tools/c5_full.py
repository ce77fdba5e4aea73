"""BASELINE config c5 at scale: a document-sharded static suffix automaton of >= 1 B tokens over the GPUs of one box,
4096 queries per step through the NVLink peer exchange (and NCCL, for comparison), plus the V = 128256 verification.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/c5_full.py --tokens-per-rank 125000000 > profiles/r02_c5_1b.json

Every rank generates and builds ONLY its own document range (synthetic corpus of samd_b200.synth.make_corpus, seeded by
rank; the V single-token documents go to the last shard, tools/gen_sam_alpaca.py:43-44 of the reference), uploads it and
drops the host copy; the shards' tokens are then exchanged over NCCL into the replicated corpus array every rank keeps
for reading drafts.  Builds are staggered (even ranks, then odd ranks) to halve the host-memory peak.
Check: for a sample of queries the merged (match length, draft) is compared with a BRUTE-FORCE search of the whole
corpus (the definition itself, SURVEY appendix A16: the longest suffix of history + token that occurs in the corpus and
the earliest end position of that occurrence) - independent of any automaton."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sam-decoding_b200"))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist

VOCAB, N_PREDICTS = 128256, 16


def brute_lookup(corpus, hist, tok):
    """(match length, earliest 1-based end position) of the longest suffix of hist + [tok] that occurs in corpus[1:]."""
    idx = (corpus == int(tok)).nonzero().flatten()
    idx = idx[idx >= 1]
    if idx.numel() == 0:
        return 0, 0
    length, best = 1, int(idx.min())
    for k in range(1, len(hist) + 1):
        idx = idx[idx - k >= 1]
        idx = idx[corpus[idx - k] == int(hist[-k])]
        if idx.numel() == 0:
            break
        length, best = k + 1, int(idx.min())
    return length, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens-per-rank", type=int, default=125_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warm", type=int, default=8)
    ap.add_argument("--brute", type=int, default=48, help="queries checked by brute force per checked step")
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from samd_b200 import dist as D, engine as E, synth
    import bench
    import psutil
    t_all = time.time()
    info = {}
    # ---- 1. this rank's documents, its automaton --------------------------------------------------------
    sam = flat_text = None
    for phase in range(2 if world > 1 else 1):
        if world == 1 or rank % 2 == phase:
            t0 = time.time()
            docs = synth.make_corpus(a.tokens_per_rank, VOCAB, 5000 + rank, singletons=(rank == world - 1))
            lens = np.fromiter((len(d) for d in docs), dtype=np.int64, count=len(docs))
            offs = np.zeros(len(docs) + 1, dtype=np.int64)
            np.cumsum(lens, out=offs[1:])
            flat = np.concatenate(docs).astype(np.int32)
            n_docs = len(docs)
            del docs
            t1 = time.time()
            sam = E.StaticSamDevice.build_flat(flat, offs, synth.EOS, False, dev)
            t2 = time.time()
            # the shard's text as the automaton indexes it (documents + the EOS appended after each): rebuilt from flat
            text = np.full(sam.n_tokens, synth.EOS, dtype=np.int32)
            appended = (flat[offs[1:] - 1] != synth.EOS).astype(np.int64)       # an EOS follows every document that lacks one
            shift = np.concatenate([[0], np.cumsum(appended)[:-1]])            # EOS tokens appended in front of document d
            assert len(flat) + int(appended.sum()) == sam.n_tokens
            dst = np.repeat(shift, lens) + np.arange(len(flat))
            text[dst] = flat
            flat_text = torch.from_numpy(text).to(dev)
            sam.drop_host()
            info = {"gen_s": t1 - t0, "build_s": t2 - t1, "tokens": sam.n_tokens, "states": sam.n_states, "edges": sam.n_edges,
                    "device_gb": sam.nbytes / 1e9, "host_rss_peak_gb": psutil.Process().memory_info().rss / 1e9, "docs": n_docs}
            del flat, text, dst
        dist.barrier()
    # ---- 2. the replicated corpus ---------------------------------------------------------------------------
    n_mine = torch.tensor([sam.n_tokens], dtype=torch.int64, device=dev)
    n_all = [torch.zeros_like(n_mine) for _ in range(world)]
    dist.all_gather(n_all, n_mine)
    sizes = [int(x.item()) for x in n_all]
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    total = int(offsets[-1])
    corpus = torch.empty(total + 1, dtype=torch.int32, device=dev)
    corpus[0] = -1
    for g in range(world):
        view = corpus[1 + offsets[g]:1 + offsets[g + 1]]
        if g == rank:
            view.copy_(flat_text)
        dist.broadcast(view, g)
    del flat_text
    sh = D.ShardedStaticSam.from_parts(sam, int(offsets[rank]), corpus, rank, world, a.queries)
    # ---- 3. queries: windows of the global corpus (70 %) and noise, EOS-free, the same on every rank -------------
    n_q, S, W = a.queries, a.steps, a.warm
    q_len = 8 * (S + W) + 1
    g = torch.Generator(device="cpu").manual_seed(5001)
    q = torch.randint(synth.FIRST_TOKEN, VOCAB, (n_q, q_len), generator=g, dtype=torch.int32)
    if rank == 0:
        qd = q.to(dev)
        for i in range(n_q):
            p = 0
            while p < q_len:
                span = int(torch.randint(4, 33, (1,), generator=g))
                span = min(span, q_len - p)
                if float(torch.rand(1, generator=g)) < 0.7:
                    off = int(torch.randint(1, total - span, (1,), generator=g))
                    qd[i, p:p + span] = corpus[off:off + span]
                p += span
        qd[qd == synth.EOS] = synth.FIRST_TOKEN
    else:
        qd = torch.empty(n_q, q_len, dtype=torch.int32, device=dev)
    dist.broadcast(qd, 0)
    q = qd.cpu().numpy()
    rng = np.random.default_rng(5002)
    counts = rng.integers(1, 9, size=(S + W, n_q)).astype(np.int32)
    ends = np.cumsum(counts, axis=0)
    begins = ends - counts
    cols = np.arange(8)[None, None, :]
    rows = np.arange(n_q)[None, :, None]
    tokens = np.where(cols < counts[:, :, None], q[rows, np.minimum(begins[:, :, None] + cols, q_len - 1)], 0).astype(np.int32)
    start = q[np.arange(n_q)[None, :], ends].astype(np.int32)
    d_tok, d_cnt, d_st = (torch.as_tensor(x).to(dev) for x in (tokens, counts, start))

    def timed(p2p, graph):
        sh.reset()
        out = (torch.empty(n_q, dtype=torch.int32, device=dev), torch.empty(n_q, N_PREDICTS, dtype=torch.int32, device=dev))
        step = lambda s: sh.lookup_draft(d_st[s], N_PREDICTS, p2p=p2p, out=out if p2p else None, tokens=d_tok[s], counts=d_cnt[s])
        for s in range(W):
            res = step(s)
        torch.cuda.synchronize()
        dist.barrier()
        gr = None
        if graph:
            cur0 = sh.cursor.clone()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for s in range(W, W + S):
                    res = step(s)
            sh.cursor.copy_(cur0)
            torch.cuda.synchronize()
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if gr is not None:
            gr.replay()
        else:
            for s in range(W, W + S):
                res = step(s)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), res[0].clone(), res[1].clone()

    ms_n, match_n, draft_n = timed(False, False)
    res = {"nccl": {"us_per_step": ms_n / S * 1e3, "queries_per_s": n_q * S / (ms_n * 1e-3)}}
    p2p_ok = sh.connect_peers()
    if p2p_ok:
        ms_p, match_p, draft_p = timed(True, False)
        ms_g, match_g, draft_g = timed(True, True)
        res["p2p"] = {"us_per_step": ms_p / S * 1e3, "queries_per_s": n_q * S / (ms_p * 1e-3), "graph_us_per_step": ms_g / S * 1e3,
                      "graph_queries_per_s": n_q * S / (ms_g * 1e-3), "peers_ok": sh.peers_ok(),
                      "identical_to_nccl": bool(torch.equal(match_p, match_n) and torch.equal(draft_p, draft_n) and
                                                torch.equal(match_g, match_n) and torch.equal(draft_g, draft_n))}
    # ---- 4. brute-force check of a sample, both exchange paths ----------------------------------------------
    check = {}
    for path in (("nccl", "p2p") if p2p_ok else ("nccl",)):
        sh.reset()
        ok, n_cmp, t0 = True, 0, time.time()
        for s in range(3):
            m, d = sh.lookup_draft(d_st[s], N_PREDICTS, p2p=(path == "p2p"), tokens=d_tok[s], counts=d_cnt[s])
            torch.cuda.synchronize()
            if rank == 0:
                m_h, d_h = m.cpu().numpy(), d.cpu().numpy()
                for i in range(0, n_q, max(1, n_q // a.brute)):
                    hist = q[i, :ends[s, i]]
                    L, e = brute_lookup(corpus, hist, start[s, i])
                    want = [int(start[s, i])] + corpus[e + 1:e + N_PREDICTS].tolist() if L else None
                    good = int(m_h[i]) == L and (L == 0 or d_h[i].tolist()[:len(want)] == want)
                    ok, n_cmp = ok and good, n_cmp + 1
        flag = torch.tensor([1 if ok else 0, n_cmp], device=dev, dtype=torch.int32)
        dist.broadcast(flag, 0)
        check[path] = {"identical": bool(flag[0].item()), "queries_checked": int(flag[1].item()), "seconds": time.time() - t0}
    # ---- 5. the verification half on this rank's own requests (V = 128256, Llama-3-8B KV) --------------------
    verify = None
    if not a.no_verify:
        class A:                                               # bench_verify's argument bag
            verify_vocab, kv_len = 128256, 8192
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        v = bench.bench_verify(A, dev, float(peaks.get("hbm_gbs", 6650.0)), vocab=128256, heads=8, kv_len=8192, name="c5", recycle=False)
        t = torch.tensor([v["us_per_step"], v["us_per_step_verify_only"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        verify = {"us_per_step_max_over_ranks": float(t[0]), "us_verify_only_max_over_ranks": float(t[1]), "frac_rank0": v["roofline"]["frac"],
                  "requests_per_step_all_ranks": 64 * world, "workload": v["workload"]}
    infos = [None] * world
    dist.all_gather_object(infos, info)
    if rank == 0:
        out = {"workload": f"c5 at scale: static SAM over {total} corpus tokens (vocab {VOCAB}) split by document over {world} GPUs, "
                           f"{n_q} queries/step advanced 1-8 tokens, draft {N_PREDICTS}; every rank generated and built only its own shard",
               "n_gpus": world, "corpus_tokens": total, "per_rank": infos, "retrieval": res, "brute_force_check": check,
               "verify_c5": verify, "mean_match": float(match_n.float().mean()), "wall_s": time.time() - t_all,
               "host_ram_total_gb": psutil.virtual_memory().total / 1e9}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
