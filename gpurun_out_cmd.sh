set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_selfplay.py -m gpu -x -q > gpurun_out/pytest_sp.log 2>&1; tail -30 gpurun_out/pytest_sp.log
