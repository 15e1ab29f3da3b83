set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_pimc.py -m gpu -x -q > gpurun_out/r02_pytest_v12.log 2>&1; tail -5 gpurun_out/r02_pytest_v12.log
python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v6.json 2> gpurun_out/r02_uct_bench_v6.err; cat gpurun_out/r02_uct_bench_v6.json; tail -3 gpurun_out/r02_uct_bench_v6.err
DOKO_CUDA_UCT_TREE_MAJOR=1 python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v6_tree_major.json 2>/dev/null; cat gpurun_out/r02_uct_bench_v6_tree_major.json
