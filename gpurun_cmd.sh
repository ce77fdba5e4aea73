python cyc.py 2>&1 | tail -12
