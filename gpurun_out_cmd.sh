set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python profiles/bench_uct.py > gpurun_out/uct_v5.json 2> gpurun_out/uct_v5.err; cat gpurun_out/uct_v5.json
python profiles/bench_kernels.py > gpurun_out/kernels_v18.json 2> gpurun_out/kernels_v18.err; tail -c 300 gpurun_out/kernels_v18.json
