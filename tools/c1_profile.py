"""Where a batch-1 step through the drop-in classes spends its time (profiling aid): cProfile of 256 update+lookup steps."""
import cProfile, pstats, sys, time
sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from samd_b200 import synth
import samd_sam_only as SO
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
prompt, steps = 4096, 256
stream = synth.copy_mix(prompt + 8 * steps + 1, 32000, 1000).astype(np.int32)
rng = np.random.default_rng(1001)
counts = rng.integers(1, 9, size=steps).astype(np.int32)
ends = prompt + np.cumsum(counts)
toks = [torch.as_tensor(stream[ends[i] - counts[i]:ends[i]]).to(dev) for i in range(steps)]
dm = SO.DraftModel(SO.SamdConfig(max_predicts=40, alpha=4.0, K=8, len_bias=5), device="cuda:0")
dm.reset()
dm.update(torch.as_tensor(stream[:prompt]).to(dev))
torch.cuda.synchronize()
def loop():
    for i in range(steps):
        dm.update(toks[i])
        dm.lookup(int(stream[ends[i]]))
t0 = time.perf_counter(); loop(); print("us/step", (time.perf_counter() - t0) / steps * 1e6)
pr = cProfile.Profile(); pr.enable(); loop(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
