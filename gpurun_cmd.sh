python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "recycle or topk or verify or kv" 2>&1 | tail -4
python tools/verify_modes.py 2>&1 | tail -4
python tools/verify_timeline.py 2>&1 | tail -5
