python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err; echo "rc=$?"; tail -c 200 gpurun_out/bench_l2.err
python -c "
import json;d=json.load(open('gpurun_out/bench_l2.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print(d['static']['queries_per_s'], d['c2_concurrent']['queries_per_s'], d['c1']['gpu_us_per_step'], d['step_kernel']['prefill_ms'])"
python tools/step_time_distribution.py 2>&1 | tail -13
