set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_pimc.py tests/test_gpu_matching.py tests/test_gpu_parity_at_size.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r02_pytest_v7.log 2>&1; tail -5 gpurun_out/r02_pytest_v7.log
python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v4_base.json 2> gpurun_out/r02_uct_bench_v4_base.err; cat gpurun_out/r02_uct_bench_v4_base.json
DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_r4.so python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v4_r4.json 2>&1; cat gpurun_out/r02_uct_bench_v4_r4.json
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v2.json 2> gpurun_out/r02_kernels_v2.err; tail -c 1200 gpurun_out/r02_kernels_v2.json
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 200 --csv --log-file gpurun_out/r02_uct_launches_v4.csv python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_tree -s 200 -c 1 -o gpurun_out/r02_uct_tree_v4 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_tree.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_rollout -s 200 -c 1 -o gpurun_out/r02_uct_rollout_v4 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_roll.log 2>&1
