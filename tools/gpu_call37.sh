set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v26.log 2>&1; tail -3 gpurun_out/r02_pytest_v26.log
timeout 120 python tools/k2_time.py 3
timeout 300 python profiles/bench_kernels.py 2> gpurun_out/r02_kernels_v12.err > gpurun_out/r02_kernels_v12.json; python -c "
import json; d=json.load(open('gpurun_out/r02_kernels_v12.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items() if a in ('sec','samples_per_s','rollouts_per_s','game_steps_per_s','iterations_per_s','decisions_per_s','hbm_frac_of_measured')})"
timeout 300 python profiles/bench_uct.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: '%.3g'%v['iterations_per_s'] for k,v in d.items()})"
