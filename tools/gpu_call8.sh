set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_state_ops.py tests/test_gpu_uct.py tests/test_gpu_selfplay.py -m gpu -x -q > gpurun_out/r02_pytest_v8.log 2>&1; tail -5 gpurun_out/r02_pytest_v8.log
python profiles/experiments/state_ops_bw.py > gpurun_out/r02_state_ops_bw_sorted.json 2>&1; cat gpurun_out/r02_state_ops_bw_sorted.json
DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_apply_nosort.so python profiles/experiments/state_ops_bw.py > gpurun_out/r02_state_ops_bw_nosort.json 2>&1; cat gpurun_out/r02_state_ops_bw_nosort.json
python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v5.json 2> gpurun_out/r02_uct_bench_v5.err; cat gpurun_out/r02_uct_bench_v5.json
python profiles/experiments/state_ops_bw.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:apply_tma -s 8 -c 1 -o gpurun_out/r02_apply_sorted python profiles/experiments/state_ops_bw.py > gpurun_out/ncu_apply.log 2>&1
