mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c 1-300 gpurun_out/bench_ref.json
