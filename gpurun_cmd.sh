mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1_final.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sam_step_kernel -s 8 -c 2 -f -o gpurun_out/prof_step_r1 python bench.py --steps 8 --warmup 3 --only-step > gpurun_out/ncu_step_full.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*verify_compact_kernel.*bool\)0.*' -s 12 -c 2 -f -o gpurun_out/prof_verify_r1 python bench.py --only-verify --kv-len 256 > gpurun_out/ncu_verify_full.log 2>&1; echo "ncu verify rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
