mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 16 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -c 400 gpurun_out/bench_n2.err; cut -c 1-600 gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; cut -c 1-500 gpurun_out/bench_ref_n2.json
