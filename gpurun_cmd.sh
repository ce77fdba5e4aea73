python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "verify or kv or recycle or topk or lossless or graph" 2>&1 | tail -4
python tools/verify_modes.py 2>&1 | tail -3
