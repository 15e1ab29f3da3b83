set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python profiles/bench_kernels.py > gpurun_out/kernels_v9.json 2> gpurun_out/kernels_v9.err; head -c 900 gpurun_out/kernels_v9.json; tail -3 gpurun_out/kernels_v9.err
