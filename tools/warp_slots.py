"""Which scheduler partition do the builder warps land on?  (profiling aid; run from the repo root on the GPU box)

A launch shaped like the c2 step (1024 CTAs x 3 warps) records every warp's %smid and %warpid; slot mod 4 is taken as the
partition.  Printed: per role (warp 0 = builder, 1 = cursor scout, 2 = redirect / short-context scouts) how the warps spread
over the four partitions, and the worst SM's number of builders on one partition."""
import sys
sys.path.insert(0, "sam-decoding_b200")
import numpy as np, torch
from samd_b200 import _cabi as K
n_blocks, warps = 1024, 3
out = torch.zeros(n_blocks * warps, 2, dtype=torch.int32, device="cuda")
K.check(K.lib().samd_debug_warp_slots(out.data_ptr(), n_blocks, warps * 32, 30000, K.stream_ptr()))
torch.cuda.synchronize()
o = out.cpu().numpy().reshape(n_blocks, warps, 2)
print("SMs used:", len(np.unique(o[:, :, 0])), " CTAs per SM: min", np.bincount(o[:, 0, 0]).min(), "max", np.bincount(o[:, 0, 0]).max())
for role in range(warps):
    part = o[:, role, 1] % 4
    print(f"warp {role}: partition histogram", np.bincount(part, minlength=4).tolist())
worst = 0
for sm in np.unique(o[:, 0, 0]):
    b = o[o[:, 0, 0] == sm][:, 0, 1] % 4
    worst = max(worst, np.bincount(b, minlength=4).max())
print("most builders of one SM on one partition:", worst)
sm0 = o[o[:, 0, 0] == o[0, 0, 0]]
print("one SM, (block, slots of its three warps):", [(int(i), sm0[i, :, 1].tolist()) for i in range(len(sm0))])
