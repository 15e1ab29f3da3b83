set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matching.py -m gpu -x -q 2>&1 | tail -1
python profiles/experiments/assign_profile_run.py > gpurun_out/assign_now3.txt 2>&1; tail -1 gpurun_out/assign_now3.txt
