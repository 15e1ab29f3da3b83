set -x
python tools/k2_time.py 6
python profiles/bench_kernels.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k,v in d.items():
    if k.startswith('K2') or k.startswith('K1') or k.startswith('K5'): print(k, v)"
