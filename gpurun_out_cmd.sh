set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_state_ops.py tests/test_gpu_matching.py tests/test_gpu_playout.py -m gpu -x -q > gpurun_out/pytest_fd.log 2>&1; tail -3 gpurun_out/pytest_fd.log
: > gpurun_out/from_deals.txt
for v in tma notma; do
  if [ $v = notma ]; then export DOKO_CUDA_NO_TMA=1; else unset DOKO_CUDA_NO_TMA; fi
  echo "variant=$v" >> gpurun_out/from_deals.txt
  timeout 300 python profiles/experiments/state_ops_bw.py >> gpurun_out/from_deals.txt 2>&1
done
tail -c 1200 gpurun_out/from_deals.txt
