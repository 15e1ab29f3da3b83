set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_matching.py tests/test_gpu_parity_at_size.py tests/test_gpu_full_size.py tests/test_gpu_cpp_env.py -m gpu -x -q > gpurun_out/r02_pytest_v4.log 2>&1; tail -5 gpurun_out/r02_pytest_v4.log
python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v2_base.json 2> gpurun_out/r02_uct_bench_v2_base.err; cat gpurun_out/r02_uct_bench_v2_base.json
DOKO_CUDA_UCT_TREE_MAJOR=1 python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v2_treemajor.json 2>&1; cat gpurun_out/r02_uct_bench_v2_treemajor.json
DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_b.so python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v2_b.json 2>&1; cat gpurun_out/r02_uct_bench_v2_b.json
DOKO_CUDA_UCT_TREE_MAJOR=1 DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_b.so python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v2_b_treemajor.json 2>&1; cat gpurun_out/r02_uct_bench_v2_b_treemajor.json
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v1.json 2> gpurun_out/r02_kernels_v1.err; tail -c 1500 gpurun_out/r02_kernels_v1.json
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 300 --csv --log-file gpurun_out/r02_uct_launches_v2.csv python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct.log 2>&1
