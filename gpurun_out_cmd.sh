set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_uct.py tests/test_replay_record.py -m gpu -x -q > gpurun_out/pytest_uct.log 2>&1; tail -25 gpurun_out/pytest_uct.log
timeout 600 python profiles/bench_kernels.py > gpurun_out/kernels_v7.json 2> gpurun_out/kernels_v7.err; tail -c 1100 gpurun_out/kernels_v7.json; tail -3 gpurun_out/kernels_v7.err
