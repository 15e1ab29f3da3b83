set -x
mkdir -p gpurun_out
DOKO_CUDA_APPLY_PIPE=1 timeout 600 python -m pytest tests/test_gpu_state_ops.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; tail -5 gpurun_out/pytest_pipe.log
: > gpurun_out/apply_tma_pipe.txt
for mb in 4 5 6; do
  export DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_applypipe$mb.so
  echo "lib=pipe$mb" >> gpurun_out/apply_tma_pipe.txt
  DOKO_CUDA_APPLY_PIPE=1 timeout 300 python profiles/experiments/state_ops_bw.py >> gpurun_out/apply_tma_pipe.txt 2>&1
done
echo "lib=nopipe" >> gpurun_out/apply_tma_pipe.txt
timeout 300 python profiles/experiments/state_ops_bw.py >> gpurun_out/apply_tma_pipe.txt 2>&1
