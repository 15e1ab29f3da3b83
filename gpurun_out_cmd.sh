set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py -m gpu -x -q > gpurun_out/pytest_host.log 2>&1; tail -3 gpurun_out/pytest_host.log
python bench.py > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err; tail -c 200 gpurun_out/bench_v16.err
