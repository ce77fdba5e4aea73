mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*bool\)1.*' --launch-skip 6 -c 1 -f -o gpurun_out/prof_verify_topk2 python bench.py --only-verify --kv-len 256 > gpurun_out/ncu_vt.log 2>&1; echo "rc=$?"
