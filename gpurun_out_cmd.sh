set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_replay_record.py tests/test_gpu_selfplay.py -m gpu -x -q > gpurun_out/pytest_replay.log 2>&1; tail -15 gpurun_out/pytest_replay.log
timeout 300 python profiles/experiments/n4_bench.py > gpurun_out/n4_bench3.txt 2>&1; tail -1 gpurun_out/n4_bench3.txt
