set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python profiles/experiments/k5_split.py > gpurun_out/k5_split.txt 2>&1; tail -1 gpurun_out/k5_split.txt
python profiles/experiments/n1_split.py > gpurun_out/n1_split3.txt 2>&1; tail -1 gpurun_out/n1_split3.txt
