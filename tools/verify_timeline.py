"""Per-warp timeline of one c4 verify launch (profiling aid; run from the repo root): when warps start, when they
finish streaming logits, when they leave."""
import sys
sys.path.insert(0, "sam-decoding_b200")
import numpy as np
import torch
from samd_b200 import _cabi as K, engine as E, synth
dev = torch.device("cuda")
B, T, V = 64, 61, 32000
for arg in sys.argv[1:]:
    if arg.startswith("B="):
        B = int(arg[2:])
move = len(sys.argv) > 1 and sys.argv[1] == "kv"
ri_np = synth.tree_retrieve_indices(synth.token_recycle_tree())
rng = np.random.default_rng(4000)
tree_tokens = rng.integers(3, V, size=(B, T)).astype(np.int32)
logits = [synth.planted_logits(B, T, V, tree_tokens, ri_np, seed=4000 + i, device=dev)[0] for i in range(4)]
ver = E.Verifier(B, T, dev)
d_tok, d_ri = torch.as_tensor(tree_tokens).to(dev), torch.as_tensor(ri_np).to(dev)
cache_len = torch.full((B,), 300, dtype=torch.int32, device=dev)
if move:
    kv_all = torch.empty(64, B, 32, 1024 if B < 64 else 2048, 128, dtype=torch.bfloat16, device=dev)
    ver.bind_kv([kv_all[i] for i in range(64)])
n_warps = 148 * 4 * 8
times = torch.zeros(n_warps, 3, dtype=torch.int64, device=dev)
out = None
for i in range(6):
    if i == 5:
        K.lib().samd_verify_set_debug_times(times.data_ptr())
    cache_len.fill_(300)
    out = ver.verify(logits[i % 4], d_tok, d_ri, cache_len=cache_len, move_kv=move, out=out)
torch.cuda.synchronize()
K.lib().samd_verify_set_debug_times(None)
t = times.cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
t = (t - t0) / 1e3
pct = lambda x: " ".join(f"{np.percentile(x, q):6.1f}" for q in (0, 10, 50, 90, 99, 100))
print(f"{len(t)} warps; us after the first warp started, percentiles 0/10/50/90/99/100")
print("start          ", pct(t[:, 0]))
print("stream end     ", pct(t[:, 1]))
busy = t[:, 1] - t[:, 0] > 5
print(f"stream end (warps that streamed, {busy.sum()})", pct(t[busy, 1]))
print("exit           ", pct(t[:, 2]))
