mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r1e.err
