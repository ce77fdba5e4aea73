"""Dependent-load latency on this GPU at the step kernel's concurrency (profiling aid; run from the repo root)."""
import sys; sys.path.insert(0, "sam-decoding_b200")
import torch
from samd_b200 import _cabi as K
dev = torch.device("cuda")
for gb, warps in ((0.05, 1024), (2.4, 1024), (16.0, 1024), (64.0, 1024), (64.0, 8192)):
    n = int(gb * 1e9 / 64)
    t = torch.randint(0, 2 ** 31 - 1, (n, 16), dtype=torch.int32, device=dev)
    sink = torch.zeros(warps, dtype=torch.int32, device=dev)
    ms = {}
    for hops in (256, 1024, 4096):
        best = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.check(K.lib().samd_debug_pointer_chase(t.data_ptr(), n, warps, hops, sink.data_ptr(), K.stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ms[hops] = best
    slope = (ms[4096] - ms[1024]) * 1e3 / 3072
    print(f"table {gb} GB, {warps} warps: {slope:.3f} us per dependent load (fixed {ms[256] * 1e3 - 256 * slope:.1f} us)")
    del t
