set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python profiles/experiments/k2_grid.py > gpurun_out/k2_now.txt 2>&1; cat gpurun_out/k2_now.txt
