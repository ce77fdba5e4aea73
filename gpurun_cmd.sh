mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 256 --warmup 16 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; wc -l gpurun_out/bench_n2.json; python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print(d['value'],d['n_gpus'],d['e2e']['value']);print(json.dumps(d['sharded_static'])[:900])"
