mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "recycle or topk or verify" 2>&1 | tail -5
python bench.py --only-verify --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
v=json.loads(sys.stdin.read())['verify']; print({k:(round(x,1) if isinstance(x,float) else x) for k,x in v.items() if k.startswith('us_')}, v['token_recycle'])"
