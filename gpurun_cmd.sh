mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1m.err
python tools/step_time_distribution.py 2>&1 | tail -11
python tools/pointer_chase.py 2>&1 | tail -8
