set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode_ipi.py -m gpu -x -q 2>&1 | tail -1
python profiles/experiments/n4_bench.py > gpurun_out/n4_bench3.txt 2>&1; tail -1 gpurun_out/n4_bench3.txt
