"""Floor of the c2 step: every request's exact sequence of record reads, replayed as bare dependent loads.

  python tools/step_floor.py [steps]        (run from the repo root, on the GPU box)

Pass A runs the real step kernel (variant 1) with the trace hook on: per request and step, the list of states whose
records the builder thread read, in order (csrc/sam_scalar.cuh ScBuilder::load).  Pass B / C start again from the same
snapshot and, before every real step (untimed, it only moves the automata and the L2 contents forward exactly as in
pass A), time a replay of that step's trace: one thread per request, four 128-bit loads per record, every address
dependent on the previous record's value - with nothing else (B), or with a second thread per request that prefetches
the whole list ahead of it (C: an ideal scout).  Then the actual kernels, both variants, with per-request SM cycles.
The step is the slowest of the 1024 requests, so everything is reported per percentile."""
import ctypes as C
import sys

sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import numpy as np, torch, bench
from samd_b200 import _cabi as K, engine as E

R, N, W = 1024, 8192, 8
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
CAP = 256
dev = torch.device("cuda")
streams, counts, tokens, start = bench.make_workload(R, N, W + S, 2000)
dyn = E.DynSamBatch(R, N + 8 * (W + S) + 16, dev)
snap = E.DynSamBatch(R, dyn.max_tokens, dev)
eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
eng.step(torch.as_tensor(streams[:, :N]).to(dev), None, None)
dt, dc, ds = (torch.as_tensor(x).to(dev) for x in (tokens, counts, start))
for s in range(W):
    eng.step(dt[s], dc[s], ds[s])
torch.cuda.synchronize()
snap.copy_from(dyn)
L = K.lib()
clk = 1.9e3        # cycles per us (the kernels are too short to leave the boost clock)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def actual(variant, trace=None):
    """per-step kernel us, per-request us [S, R]"""
    L.samd_step_set_variant(variant)
    dyn.copy_from(snap); torch.cuda.synchronize()
    cyc = torch.zeros(10, R, dtype=torch.int64, device=dev)
    L.samd_step_set_debug_cycles(cyc.data_ptr())
    us, per = [], []
    for i, s in enumerate(range(W, W + S)):
        if trace is not None:
            L.samd_step_set_trace(trace[i].data_ptr(), CAP)
        us.append(timed(lambda: eng.step(dt[s], dc[s], ds[s])))
        per.append(cyc[0].cpu().numpy() / clk)
    L.samd_step_set_debug_cycles(None); L.samd_step_set_trace(None, 0); L.samd_step_set_variant(1)
    return np.array(us), np.stack(per)


def plain(variant):
    """per-step us of the production build of the kernel (no cycle counters)"""
    L.samd_step_set_variant(variant)
    dyn.copy_from(snap); torch.cuda.synchronize()
    us = [timed(lambda: eng.step(dt[s], dc[s], ds[s])) for s in range(W, W + S)]
    L.samd_step_set_variant(1)
    return np.array(us)


def replay(trace, with_scout):
    dyn.copy_from(snap); torch.cuda.synchronize()
    cyc = torch.zeros(R, dtype=torch.int64, device=dev)
    us, per = [], []
    for i, s in enumerate(range(W, W + S)):
        us.append(timed(lambda: K.check(L.samd_debug_replay_trace(dyn.handle, trace[i].data_ptr(), CAP, with_scout, cyc.data_ptr(), K.stream_ptr()))))
        per.append(cyc.cpu().numpy() / clk)
        eng.step(dt[s], dc[s], ds[s])          # untimed: move the automata (and L2) forward as the real run does
    torch.cuda.synchronize()
    return np.array(us), np.stack(per)


def line(name, us, per):
    p = lambda q: np.percentile(per, q)
    print(f"{name:34s} kernel us/step mean {us.mean():5.1f} | per request us: mean {per.mean():5.2f} p50 {p(50):5.2f} p90 {p(90):5.2f} "
          f"p99 {p(99):5.2f} p99.9 {p(99.9):5.2f} | slowest request of a step: mean {per.max(1).mean():5.1f}")


trace = torch.zeros(S, R, CAP, dtype=torch.int32, device=dev)
us1, per1 = actual(1, trace)
n_loads = trace[:, :, 0].cpu().numpy()
print(f"c2: {R} requests x {N}-token prompts, {S} steps; builder record reads per request-step: mean {n_loads.mean():.1f} "
      f"p50 {np.percentile(n_loads, 50):.0f} p99 {np.percentile(n_loads, 99):.0f} max {n_loads.max()} (trace capacity {CAP - 1})")
line("actual, one thread per request", us1, per1)
us0, per0 = actual(0)
line("actual, warp-cooperative (round 1)", us0, per0)
print(f"production builds (no counters): variant 1 {plain(1).mean():.1f} us/step, variant 0 {plain(0).mean():.1f} us/step")
usb, perb = replay(trace, 0)
line("floor: bare dependent loads", usb, perb)
usc, perc = replay(trace, 1)
line("floor: + ideal prefetcher", usc, perc)
slow = per1 > np.percentile(per1, 99.5)
print(f"slowest 0.5% of request-steps: actual {per1[slow].mean():.1f} us, {n_loads[slow].mean():.0f} record reads, "
      f"bare-load floor {perb[slow].mean():.1f} us, with ideal prefetcher {perc[slow].mean():.1f} us")
k = counts[W:W + S]
for kk in (1, 4, 8):
    m = k == kk
    print(f"  k={kk}: reads {n_loads[m].mean():5.1f}  actual {per1[m].mean():5.2f} us  floor {perb[m].mean():5.2f}  floor+prefetch {perc[m].mean():5.2f}")
