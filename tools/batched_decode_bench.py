"""Steps per second of the batched decode loop at B = 64 on a tiny random-init Llama: the CUDA-graphed
loop (one host sync per 4 steps) against the same body launched eagerly and against plain greedy decoding.
    python tools/batched_decode_bench.py        (GPU box, from the repo root; prints one JSON line)"""
import json
import sys
import time

sys.path.insert(0, "sam-decoding_b200")
import torch
from transformers import LlamaConfig, LlamaForCausalLM
from samd_b200 import synth
from samd_b200.batched import BatchedSamdDecoder

torch.manual_seed(0)
cfg = LlamaConfig(vocab_size=96, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                  num_key_value_heads=4, max_position_embeddings=2048)
lm = LlamaForCausalLM(cfg).cuda().to(torch.float16).eval()
B, n_new = 64, 128
prompts = [synth.copy_mix(192 + (i % 5) * 16, 96, 900 + i, p_copy=0.7).tolist() for i in range(B)]
dec = BatchedSamdDecoder(lm, B, 1024, n_predicts=8, len_bias=5, len_threshold=3, dtype=torch.float16)
res = {}
for name, kw in (("graph", dict(graph=True)), ("eager", dict(graph=False))):
    dec.generate(prompts, n_new, **kw)                       # warm-up (graph capture, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, st = dec.generate(prompts, n_new, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[name] = {"steps": st["steps_run"], "host_syncs": st["host_syncs"], "graphed": st["graphed"], "graphs": st["graphs"],
                 "seconds": dt, "steps_per_s": st["steps_run"] / dt, "tokens_per_s": sum(len(o) for o in out) / dt,
                 "mean_accept": sum(map(sum, st["accept_lengths"])) / max(1, st["steps"] * B)}
    res[name + "_tokens"] = out
assert res.pop("graph_tokens") == res.pop("eager_tokens")
print(json.dumps({"workload": f"batched decode, B={B}, tiny Llama (2 layers, d=128, V=96) fp16, {n_new} new tokens, n_predicts 8; "
                              "`seconds` = one generate() call incl. prefill; the graphs were captured by the warm-up call and are replayed", **res}))
