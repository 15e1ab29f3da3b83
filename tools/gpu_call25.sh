set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matching.py tests/test_gpu_full_size.py tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_v20.log 2>&1; tail -3 gpurun_out/r02_pytest_v20.log
python profiles/bench_kernels.py 2> gpurun_out/r02_kernels_v7.err > gpurun_out/r02_kernels_v7.json; python -c "
import json; d=json.load(open('gpurun_out/r02_kernels_v7.json'))
for k,v in d.items():
    if k.startswith('K3') or k.startswith('K4'): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items()})"
python bench.py > gpurun_out/r02_bench_v3_1gpu.json 2> gpurun_out/r02_bench_v3_1gpu.err; tail -c 300 gpurun_out/r02_bench_v3_1gpu.err
