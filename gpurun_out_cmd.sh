set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_state_ops.py tests/test_gpu_selfplay.py -m gpu -x -q > gpurun_out/pytest_apply.log 2>&1; tail -3 gpurun_out/pytest_apply.log
: > gpurun_out/apply_pipe.txt
for mb in 4 5 6; do
  export DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_applypipe$mb.so
  echo "lib=pipe$mb" >> gpurun_out/apply_pipe.txt
  python profiles/experiments/state_ops_bw.py >> gpurun_out/apply_pipe.txt 2>&1
done
cat gpurun_out/apply_pipe.txt | cut -c1-400
