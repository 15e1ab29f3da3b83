set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matching.py tests/test_gpu_pimc.py tests/test_gpu_full_size.py tests/test_gpu_playout.py -m gpu -x -q > gpurun_out/pytest_split.log 2>&1; tail -5 gpurun_out/pytest_split.log
echo split > gpurun_out/k4_sizes.txt; timeout 600 python profiles/experiments/k4_sizes.py >> gpurun_out/k4_sizes.txt 2>&1
echo nosplit >> gpurun_out/k4_sizes.txt; DOKO_CUDA_NO_SPLIT=1 timeout 600 python profiles/experiments/k4_sizes.py >> gpurun_out/k4_sizes.txt 2>&1
