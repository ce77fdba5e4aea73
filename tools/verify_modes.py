"""Times the c4 verify + KV-compaction launch under the tuning hooks (profiling aid; run from the repo root)."""
import sys, types
sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import torch
import bench
from samd_b200 import _cabi as K
dev = torch.device("cuda")
a = types.SimpleNamespace(verify_vocab=int(sys.argv[1]) if len(sys.argv) > 1 else 32000, kv_len=2048)
for chunk in (0, 16000, 10672, 8000, 5336, 4000):
    K.lib().samd_verify_set_chunk(chunk)
    r = bench.bench_verify(a, dev, 6458.1)
    print(f"chunk {chunk}: full {r['us_per_step']:.1f} us ({r['roofline']['frac']:.3f}), verify-only "
          f"{r['us_per_step_verify_only']:.1f} us ({r['roofline_verify_only']['frac']:.3f}), kv alone {r['us_kv_compact_standalone']:.1f} us",
          flush=True)
    torch.cuda.empty_cache()
