set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_replay_record.py -x -q 2>&1 | tail -3
python profiles/experiments/n4_bench.py > gpurun_out/n4_bench2.txt 2>&1; tail -1 gpurun_out/n4_bench2.txt
