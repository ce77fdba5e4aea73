python scout_ab.py 2>&1 | tail -5
