set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matching.py tests/test_gpu_pimc.py tests/test_gpu_uct.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -1
python profiles/experiments/det_profile_run.py > gpurun_out/det_now3.txt 2>&1; tail -1 gpurun_out/det_now3.txt
python profiles/bench_kernels.py > gpurun_out/kernels_v25.json 2> gpurun_out/kernels_v25.err
