"""Ceiling of c4's KV compaction: scattered row moves as a function of the granule size.

  python tools/row_move_ceiling.py          (run from the repo root, on the GPU box)

c4 moves ~160 accepted rows per step; in the reference's cache layout ([batch, heads, positions, head_dim] per layer and
K / V, samd/cache.py:118-133) a row is 64 x 32 separate 256-byte granules, 534 KB apart - 84 MB per step as 328 k reads and
328 k writes of 256 bytes scattered over a 70 GB cache.  This tool moves the SAME number of bytes with a bare copy kernel
(samd_debug_granule_copy: offsets from a table, four 16-byte units in flight per lane, nothing else) at granule sizes
256 B ... 64 KB, the source granules spread uniformly over a 64 GB buffer and each destination 10 KB in front of its
source (the tree window), and - as the other end - one contiguous copy of the same size."""
import sys
sys.path.insert(0, "sam-decoding_b200")
import numpy as np, torch
from samd_b200 import _cabi as K
dev = torch.device("cuda")
L = K.lib()
free_b, _ = torch.cuda.mem_get_info(dev)
BUF = min(64 << 30, int(free_b * 0.8)) // (1 << 20) * (1 << 20)
buf = torch.empty(BUF, dtype=torch.uint8, device=dev)
MOVE = 42 << 20                                      # bytes read (and as many written) per launch, as c4
rng = np.random.default_rng(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(granule, blocks, reps=12):
    """Cells of 32 KB (4 granules for the large ones), one move per cell, cells drawn without repetition over the whole
    buffer: source in the middle of the cell, destination 10 KB (2 granules) in front of it."""
    n = MOVE // granule
    cell = 32768 if granule <= 4096 else 4 * granule
    cells = rng.permutation(BUF // cell)[:n].astype(np.int64)
    src = torch.as_tensor(cells * cell + (16384 if granule <= 4096 else 2 * granule)).to(dev)
    dst = src - (10240 if granule <= 4096 else 2 * granule)
    ts = []
    for _ in range(reps):
        flush.fill_(1)                                           # L2 does not hold the previous launch's lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.check(L.samd_debug_granule_copy(buf.data_ptr(), src.data_ptr(), dst.data_ptr(), n, granule, blocks, K.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))


print(f"buffer {BUF / 2**30:.0f} GB, {MOVE >> 20} MB read + {MOVE >> 20} MB written per launch, L2 flushed between launches")
for granule in (256, 512, 1024, 4096, 16384, 65536):
    best = min((run(granule, blocks), blocks) for blocks in (148 * 4, 148 * 8, 148 * 16))
    print(f"granule {granule:6d} B: {best[0]:6.1f} us  = {2 * MOVE / best[0] / 1e6:5.2f} TB/s (grid {best[1]} x 256)")
a = torch.empty(MOVE, dtype=torch.uint8, device=dev); b = torch.empty(MOVE, dtype=torch.uint8, device=dev)
ts = []
for _ in range(12):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
t = float(np.median(ts[2:]))
print(f"contiguous copy of {MOVE >> 20} MB (torch): {t:6.1f} us  = {2 * MOVE / t / 1e6:5.2f} TB/s")
