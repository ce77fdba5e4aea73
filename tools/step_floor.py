"""Floor of the c2 step: every request's exact sequence of record reads, replayed as bare dependent loads.

  python tools/step_floor.py [steps]        (run from the repo root, on the GPU box)

Pass A runs the real step kernel (variant 1, profiling build) with the trace hook on: per request and step, the list of
states whose records the builder thread read, in order (csrc/sam_scalar.cuh ScBuilderT::load), the cycles it waited for
those loads and how many of them came back in < 120 / < 500 / < 1100 / more cycles (L1 / L2 / DRAM / slower).
Pass B / C start again from the same snapshot and, before every real step (untimed, it only moves the automata and the L2
contents forward exactly as in pass A), run a replay of that step's trace: one thread per request, four 128-bit loads
per record, every address dependent on the previous record's value - with nothing else (B), or with a second thread per
request that prefetches the whole list ahead of it (C: an ideal scout).  The step is the slowest of the 1024 requests, so
everything is reported per percentile; a step's span is max(end) - min(start) of %globaltimer over its requests."""
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
clk = float(sys.argv[2]) if len(sys.argv) > 2 else 1.965e3        # SM cycles per us (the kernels run at the boost clock)


def actual(variant, trace=None, scouts=2):
    """[S, 12, R] profiling rows of every step"""
    L.samd_step_set_variant(variant); L.samd_step_set_scouts(scouts)
    dyn.copy_from(snap); torch.cuda.synchronize()
    cyc = torch.zeros(16, R, dtype=torch.int64, device=dev)
    L.samd_step_set_debug_cycles(cyc.data_ptr())
    rows = []
    for i, s in enumerate(range(W, W + S)):
        if trace is not None:
            L.samd_step_set_trace(trace[i].data_ptr(), CAP)
        eng.step(dt[s], dc[s], ds[s]); torch.cuda.synchronize()
        rows.append(cyc.cpu().numpy().copy())
    L.samd_step_set_debug_cycles(None); L.samd_step_set_trace(None, 0); L.samd_step_set_variant(1); L.samd_step_set_scouts(2)
    return np.stack(rows).astype(float)


def graph_us(variant, scouts=2):
    """production build, the S steps as one CUDA graph: us per step"""
    L.samd_step_set_variant(variant); L.samd_step_set_scouts(scouts)
    dyn.copy_from(snap); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s in range(W, W + S):
            eng.step(dt[s], dc[s], ds[s])
    g.replay(); torch.cuda.synchronize()
    dyn.copy_from(snap); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    L.samd_step_set_variant(1); L.samd_step_set_scouts(2)
    return e0.elapsed_time(e1) * 1e3 / S


def replay(trace, with_scout):
    dyn.copy_from(snap); torch.cuda.synchronize()
    cyc = torch.zeros(3, R, dtype=torch.int64, device=dev)
    rows = []
    for i, s in enumerate(range(W, W + S)):
        K.check(L.samd_debug_replay_trace(dyn.handle, trace[i].data_ptr(), CAP, with_scout, cyc.data_ptr(), K.stream_ptr()))
        torch.cuda.synchronize()
        rows.append(cyc.cpu().numpy().copy())
        eng.step(dt[s], dc[s], ds[s])          # untimed: move the automata (and L2) forward as the real run does
    torch.cuda.synchronize()
    return np.stack(rows).astype(float)


def line(name, per, span):
    p = lambda q: np.percentile(per, q)
    print(f"{name:36s} per request us: mean {per.mean():5.2f} p50 {p(50):5.2f} p90 {p(90):5.2f} p99 {p(99):5.2f} p99.9 {p(99.9):5.2f} | "
          f"slowest request of a step {per.max(1).mean():5.1f} | step span (globaltimer) {span.mean():5.1f} us")


print(f"c2: {R} requests x {N}-token prompts, {S} steps, clock {clk / 1e3:.3f} GHz assumed")
print("graph-replayed production kernels, us/step: variant 1 scouts 2/1/0: %.1f %.1f %.1f | variant 0 scouts 2/1/0: %.1f %.1f %.1f" %
      (graph_us(1, 2), graph_us(1, 1), graph_us(1, 0), graph_us(0, 2), graph_us(0, 1), graph_us(0, 0)))
trace = torch.zeros(S, R, CAP, dtype=torch.int32, device=dev)
a1 = actual(1, trace)
n_loads = trace[:, :, 0].cpu().numpy()
per1 = a1[:, 0] / clk
span1 = (a1[:, 11].max(1) - a1[:, 10].min(1)) / 1e3
print(f"builder record reads per request-step: mean {n_loads.mean():.1f} p50 {np.percentile(n_loads, 50):.0f} p99 {np.percentile(n_loads, 99):.0f} "
      f"max {n_loads.max()} (trace capacity {CAP - 1})")
line("actual, one thread per request", per1, span1)
a0 = actual(0)
line("actual, warp-cooperative (round 1)", a0[:, 0] / clk, np.zeros(S))
a1n = actual(1, None, scouts=0)
line("actual, one thread, no scouts", a1n[:, 0] / clk, (a1n[:, 11].max(1) - a1n[:, 10].min(1)) / 1e3)


def breakdown(a, name, sel=None):
    f = (lambda x: x[sel].mean()) if sel is not None else (lambda x: x.mean())
    tot, ld, up, lk = f(a[:, 0]), f(a[:, 1]), f(a[:, 2]), f(a[:, 3])
    c = [f(a[:, 4 + i]) for i in range(4)]
    print(f"  {name}: total {tot / clk:5.2f} us = update {up / clk:5.2f} + lookup/draft {lk / clk:5.2f}; waiting for record loads {ld / clk:5.2f} us "
          f"({sum(c):.1f} loads: {c[0]:.1f} <120 cyc, {c[1]:.1f} <500, {c[2]:.1f} <1100, {c[3]:.1f} slower; mean {ld / max(sum(c), 1e-9):.0f} cyc); "
          f"overflow probes {f(a[:, 8]) / clk:5.2f} us ({f(a[:, 9]):.1f}); before the first token {f(a[:, 14]) / clk:5.2f}, cursor walks "
          f"{f(a[:, 12]) / clk:5.2f}, redirect walks {f(a[:, 13]) / clk:5.2f} ({f(a[:, 15]):.1f} records read), rest of the appends {(up - f(a[:, 14]) - f(a[:, 12]) - f(a[:, 13])) / clk:5.2f} us")


slow = per1 > np.percentile(per1, 99.5)
breakdown(a1, "with scouts, all requests   ")
breakdown(a1, "with scouts, slowest 0.5%   ", slow)
breakdown(a1n, "no scouts, all requests     ")
breakdown(a1n, "no scouts, slowest 0.5%     ", (a1n[:, 0] / clk) > np.percentile(a1n[:, 0] / clk, 99.5))
rb = replay(trace, 0)
line("floor: bare dependent loads", rb[:, 0] / clk, (rb[:, 2].max(1) - rb[:, 1].min(1)) / 1e3)
rc = replay(trace, 1)
line("floor: + ideal prefetcher", rc[:, 0] / clk, (rc[:, 2].max(1) - rc[:, 1].min(1)) / 1e3)
print(f"slowest 0.5% of request-steps: actual {per1[slow].mean():.1f} us, {n_loads[slow].mean():.0f} record reads, "
      f"bare-load floor {(rb[:, 0] / clk)[slow].mean():.1f} us ({(rb[:, 0])[slow].mean() / n_loads[slow].mean():.0f} cyc per load), "
      f"with ideal prefetcher {(rc[:, 0] / clk)[slow].mean():.1f} us")
k = counts[W:W + S]
for kk in (1, 4, 8):
    m = k == kk
    print(f"  k={kk}: reads {n_loads[m].mean():5.1f}  actual {per1[m].mean():5.2f} us  floor {(rb[:, 0] / clk)[m].mean():5.2f}  floor+prefetch {(rc[:, 0] / clk)[m].mean():5.2f}")
