set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_v17.log 2>&1; tail -3 gpurun_out/r02_pytest_v17.log
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v6.json 2> gpurun_out/r02_kernels_v6.err; tail -2 gpurun_out/r02_kernels_v6.err
python -c "
import json; d=json.load(open('gpurun_out/r02_kernels_v6.json'))
for k,v in d.items(): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)"
python profiles/experiments/playout_states.py 2>&1 | tail -3
