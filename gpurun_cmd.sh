timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/p2p_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -40
