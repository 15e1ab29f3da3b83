set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python profiles/sanitize_smoke.py > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python profiles/bench_kernels.py > gpurun_out/kernels_v15.json 2> gpurun_out/kernels_v15.err; head -c 1300 gpurun_out/kernels_v15.json | tail -c 700; tail -3 gpurun_out/kernels_v15.err
