mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-extras > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1o.err; cat gpurun_out/bench_r1o.json | cut -c 1-900
