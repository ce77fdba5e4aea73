python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_cases" 2>&1 | tail -25
