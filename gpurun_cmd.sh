python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "dyn or c2_full or c3 or step_host or static or select or loop" 2>&1 | tail -3
python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'],d['e2e']['value'])"
python tools/step_time_distribution.py 2>&1 | grep "fit us\|per-step max"
