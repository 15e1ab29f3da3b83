set -x
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python -m pytest tests/test_gpu_selfplay.py tests/test_replay_record.py -m gpu -x -q 2>&1 | tail -1; done
python profiles/experiments/n1_split.py > gpurun_out/n1_split6.txt 2>&1; tail -1 gpurun_out/n1_split6.txt
