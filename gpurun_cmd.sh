timeout 300 python tools/verify_timeline.py kv B=1 2>&1 | tail -5
timeout 300 python tools/verify_timeline.py B=1 2>&1 | tail -5
