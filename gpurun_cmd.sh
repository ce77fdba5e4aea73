mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r1h.err
