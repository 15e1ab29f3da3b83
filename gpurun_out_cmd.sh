set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python profiles/bench_uct.py > gpurun_out/uct_v4.json 2> gpurun_out/uct_v4.err; cat gpurun_out/uct_v4.json; tail -3 gpurun_out/uct_v4.err
timeout 600 python profiles/bench_kernels.py > gpurun_out/kernels_v13.json 2> gpurun_out/kernels_v13.err; tail -c 900 gpurun_out/kernels_v13.json; tail -3 gpurun_out/kernels_v13.err
