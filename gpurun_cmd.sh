mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1n.err
python bench.py --only-verify --no-cpu --verify-vocab 128256 2>/dev/null | tail -1 > gpurun_out/verify_c5.json
python tools/verify_timeline.py kv 2>&1 | tail -6
