"""c2 lookup-only launches (no appends) for the sectors-per-query figure of SURVEY 8(d); run under
ncu --metrics lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct
(profiling aid; run from the repo root)."""
import sys
sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from samd_b200 import _cabi as K, engine as E
dev = torch.device("cuda")
R, N, S = 1024, 8192, 8
streams, counts, tokens, start = bench.make_workload(R, N, S, 2000)
dyn = E.DynSamBatch(R, N + 8 * S + 16, dev)
eng = E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=bench.N_PREDICTS, len_bias=bench.LEN_BIAS, len_threshold=bench.LEN_THRESHOLD)
eng.step(torch.as_tensor(streams[:, :N]).to(dev), None, None)
d_tokens, d_counts, d_start = (torch.as_tensor(x).to(dev) for x in (tokens, counts, start))
for s in range(S):                                   # append, then a separate lookup-only launch
    eng.step(d_tokens[s], d_counts[s], None)
    st0 = dyn.stats()
    eng.step(None, None, d_start[s])
    st1 = dyn.stats()
torch.cuda.synchronize()
print("lookup probes per query (last step):", (st1["lookup_probes"] - st0["lookup_probes"]) / R,
      "mean match", float(eng.match_dyn.float().mean()))
