set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_cpp_env.py tests/test_gpu_matching.py tests/test_replay_record.py -m gpu -x -q > gpurun_out/r02_pytest_v2.log 2>&1; tail -15 gpurun_out/r02_pytest_v2.log
for v in "" uct_a uct_b uct_c; do
  if [ -n "$v" ]; then export DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_$v.so; fi
  python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v1_$v.json 2> gpurun_out/r02_uct_bench_v1_$v.err; cat gpurun_out/r02_uct_bench_v1_$v.json; tail -3 gpurun_out/r02_uct_bench_v1_$v.err
done
unset DOKO_CUDA_LIB
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 300 --csv --log-file gpurun_out/r02_uct_launches_v1.csv python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct.log 2>&1
tail -5 gpurun_out/ncu_uct.log
