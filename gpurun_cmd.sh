mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "verify or kv or lossless or loop or eval_posterior or cache" 2>&1 | tail -3
python bench.py --only-verify > gpurun_out/verify_c.json 2> gpurun_out/verify_c.err; echo "rc=$?"; tail -c 400 gpurun_out/verify_c.err
python bench.py --only-verify --verify-vocab 128256 --kv-len 512 > gpurun_out/verify_c5.json 2>> gpurun_out/verify_c.err; echo "rc=$?"
