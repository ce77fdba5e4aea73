timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'],d['e2e']['value'],d['step_kernel']['prefill_ms'])"
timeout 300 python tools/step_time_distribution.py 2>&1 | tail -8
