"""Host-to-host latency of one c2 step under different staging strategies (profiling aid; run from the repo root)."""
import sys, time
sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from samd_b200 import _cabi as K, engine as E
dev = torch.device("cuda")
R, N, S, W = 1024, 8192, 256, 16
streams, counts, tokens, start = bench.make_workload(R, N, S + W, 2000)
dyn = E.DynSamBatch(R, N + 8 * (S + W) + 16, dev)
snap = E.DynSamBatch(R, N + 8 * (S + W) + 16, dev)
eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=bench.N_PREDICTS, len_bias=bench.LEN_BIAS, len_threshold=bench.LEN_THRESHOLD)
eng.step(torch.as_tensor(streams[:, :N]).to(dev), None, None)
torch.cuda.synchronize()
snap.copy_from(dyn)
inp, res = eng.host_buffers(8)
h_in = torch.empty(W + S, inp.numel(), dtype=torch.int32).pin_memory()
h_in[:, :R] = torch.as_tensor(counts)
h_in[:, R:2 * R] = torch.as_tensor(start)
h_in[:, 2 * R:] = torch.as_tensor(tokens).reshape(W + S, R * 8)
dev_in, k = eng._io
B = R


def graph_of(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def v_copy():
    dev_in.copy_(inp, non_blocking=True)
    eng.step(dev_in[2 * B:].view(B, k), dev_in[:B], dev_in[B:2 * B])
    res.copy_(eng.out_buf, non_blocking=True)


def v_zin():
    eng.step(inp[2 * B:].view(B, k), inp[:B], inp[B:2 * B])
    res.copy_(eng.out_buf, non_blocking=True)


def v_zboth():
    eng.step(inp[2 * B:].view(B, k), inp[:B], inp[B:2 * B], out_buf=res)


def v_kernel_only():
    eng.step(dev_in[2 * B:].view(B, k), dev_in[:B], dev_in[B:2 * B])


stream = torch.cuda.current_stream(dev)
for name, fn in (("copy in + copy out", v_copy), ("zero-copy in + copy out", v_zin), ("zero-copy both", v_zboth),
                 ("kernel only (no I/O)", v_kernel_only)):
    inp.zero_()
    g = graph_of(fn)
    for sync in ("stream", "event-spin"):
        dyn.copy_from(snap)
        torch.cuda.synchronize()
        ev = torch.cuda.Event()
        t0 = time.perf_counter()
        for s in range(W, W + S):
            inp.copy_(h_in[s])
            g.replay()
            if sync == "stream":
                stream.synchronize()
            else:
                ev.record()
                while not ev.query():
                    pass
        dt = (time.perf_counter() - t0) / S * 1e6
        print(f"{name:28s} sync={sync:10s}: {dt:6.1f} us per step", flush=True)
