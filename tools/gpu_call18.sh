set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matching.py tests/test_gpu_pimc.py tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_v16.log 2>&1; tail -3 gpurun_out/r02_pytest_v16.log
python profiles/experiments/det_profile_run.py > gpurun_out/r02_k3_plain_v3.json 2>&1; cat gpurun_out/r02_k3_plain_v3.json
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v5.json 2> gpurun_out/r02_kernels_v5.err; tail -2 gpurun_out/r02_kernels_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_kernels_v5.json'))
for k,v in d.items():
    if k.startswith('K3') or k.startswith('K4') or k.startswith('N2'): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items()})"
