set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pimc.py -m gpu -x -q > gpurun_out/pytest_pimc.log 2>&1; tail -15 gpurun_out/pytest_pimc.log
python profiles/bench_kernels.py > gpurun_out/kernels_v5.json 2> gpurun_out/kernels_v5.err; tail -c 600 gpurun_out/kernels_v5.json; tail -3 gpurun_out/kernels_v5.err
