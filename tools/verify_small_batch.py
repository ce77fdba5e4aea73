"""Latency of one verify + KV-compaction launch at small batch sizes, graph-replayed (profiling aid; run from the
repo root): the reference's own use is batch 1."""
import sys
sys.path.insert(0, "sam-decoding_b200")
import numpy as np
import torch
from samd_b200 import engine as E, synth
dev = torch.device("cuda")
T, V = 61, 32000
ri_np = synth.tree_retrieve_indices(synth.token_recycle_tree())
d_ri = torch.as_tensor(ri_np).to(dev)
for B in (1, 2, 4, 8, 16):
    rng = np.random.default_rng(B)
    tok = rng.integers(3, V, size=(B, T)).astype(np.int32)
    logits = [synth.planted_logits(B, T, V, tok, ri_np, seed=10 + i, device=dev)[0] for i in range(4)]
    kv = torch.zeros(64, B, 32, 1024, 128, dtype=torch.bfloat16, device=dev)
    ver = E.Verifier(B, T, dev)
    ver.bind_kv([kv[i] for i in range(64)])
    d_tok = torch.as_tensor(tok).to(dev)
    cache_len = torch.full((B,), 100, dtype=torch.int32, device=dev)
    table = E.RecycleTable(synth.token_recycle_tree(), V, dev)
    for rec in (None, table):
        outs = [ver.verify(logits[i], d_tok, d_ri, cache_len=cache_len, recycle=rec) for i in range(4)]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(16):
                ver.verify(logits[i % 4], d_tok, d_ri, cache_len=cache_len, recycle=rec, out=outs[i % 4])
        ts = []
        for rep in range(5):
            cache_len.fill_(100)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 16 * 1e3)
        print(f"B={B:2d} {'verify+KV+top8' if rec is not None else 'verify+KV     '}: {np.median(ts):6.1f} us per launch "
              f"({B * T * V * 2 / np.median(ts) / 1e3:.0f} GB/s of logits)", flush=True)
