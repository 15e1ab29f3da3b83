set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_selfplay.py tests/test_replay_record.py tests/test_gpu_state_ops.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python profiles/experiments/n1_split.py > gpurun_out/n1_split2.txt 2>&1; tail -1 gpurun_out/n1_split2.txt
