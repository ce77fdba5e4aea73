"""Document-sharded static SAM over N GPUs: NCCL all-reduce-max against the NVLink peer exchange, with an identity
check (profiling aid / multi-GPU check).  Run from the repo root:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_check.py"""
import json, os, sys, types
sys.path.insert(0, "sam-decoding_b200"); sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
res = bench.bench_sharded_static(types.SimpleNamespace(), dev, rank, world,
                                 tokens_per_shard=int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000)
if rank == 0:
    print(json.dumps(res, indent=1))
dist.barrier()
dist.destroy_process_group()
