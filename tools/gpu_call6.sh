set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_pimc.py -m gpu -x -q > gpurun_out/r02_pytest_v6.log 2>&1; tail -5 gpurun_out/r02_pytest_v6.log
python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v3_base.json 2> gpurun_out/r02_uct_bench_v3_base.err; cat gpurun_out/r02_uct_bench_v3_base.json
DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_r4.so python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v3_r4.json 2>&1; cat gpurun_out/r02_uct_bench_v3_r4.json
DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_t6.so python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v3_t6.json 2>&1; cat gpurun_out/r02_uct_bench_v3_t6.json
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 200 --csv --log-file gpurun_out/r02_uct_launches_v3.csv python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_tree -s 200 -c 1 -o gpurun_out/r02_uct_tree_v3 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_tree.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_rollout -s 200 -c 1 -o gpurun_out/r02_uct_rollout_v3 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_roll.log 2>&1
