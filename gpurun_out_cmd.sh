set -x
mkdir -p gpurun_out
for rep in 1 2 3; do for l in 64_16_0 128_8_1; do DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_uct_$l.so python profiles/experiments/uct_cfg.py; done; done > gpurun_out/uct_cfg2.txt 2>&1; cat gpurun_out/uct_cfg2.txt | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['lib'], [round(d[k]['Miter_per_s']) for k in d if k != 'lib'])
"
