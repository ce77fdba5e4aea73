mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_full.err
python -c "
import json;d=json.load(open('gpurun_out/bench_full.json'));print(d['value'],d['e2e']);print(d['c2_concurrent'])"
