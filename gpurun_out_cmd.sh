set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_state_ops.py tests/test_gpu_full_size.py tests/test_gpu_selfplay.py -m gpu -x -q > gpurun_out/pytest_ng.log 2>&1; tail -3 gpurun_out/pytest_ng.log
: > gpurun_out/new_games_tma.txt
for v in tma notma tma notma; do
  if [ $v = notma ]; then export DOKO_CUDA_NO_TMA=1; else unset DOKO_CUDA_NO_TMA; fi
  echo "variant=$v" >> gpurun_out/new_games_tma.txt
  timeout 300 python profiles/experiments/state_ops_bw.py >> gpurun_out/new_games_tma.txt 2>&1
done
