set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python profiles/bench_kernels.py > gpurun_out/kernels_v4.json 2> gpurun_out/kernels_v4.err; tail -c 1500 gpurun_out/kernels_v4.json
