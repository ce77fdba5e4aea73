python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3_static" 2>&1 | tail -15
