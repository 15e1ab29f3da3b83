set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matching.py tests/test_gpu_pimc.py tests/test_gpu_uct.py tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_v14.log 2>&1; tail -3 gpurun_out/r02_pytest_v14.log
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v4.json 2> gpurun_out/r02_kernels_v4.err; tail -2 gpurun_out/r02_kernels_v4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_kernels_v4.json'))
for k,v in d.items(): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)"
