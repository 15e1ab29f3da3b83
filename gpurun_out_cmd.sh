set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_state_ops.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; tail -15 gpurun_out/pytest_ab.log
