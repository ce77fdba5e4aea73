mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --only-verify > gpurun_out/verify_c.json 2> gpurun_out/verify_c.err; echo "rc=$?"; tail -c 300 gpurun_out/verify_c.err
