set -x
mkdir -p gpurun_out
: > gpurun_out/apply_idx.txt
for mb in 8 10 12 default; do
  if [ $mb = default ]; then unset DOKO_CUDA_LIB; else export DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_applyidx$mb.so; fi
  echo "lib=idx$mb" >> gpurun_out/apply_idx.txt
  timeout 300 python profiles/experiments/state_ops_bw.py >> gpurun_out/apply_idx.txt 2>&1
done
