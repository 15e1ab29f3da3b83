set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python profiles/bench_kernels.py > gpurun_out/kernels_v30.json 2> gpurun_out/kernels_v30.err
