python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "batched" 2>&1 | tail -30
