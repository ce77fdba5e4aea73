mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r1c.err
