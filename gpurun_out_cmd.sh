set -x
mkdir -p gpurun_out
: > gpurun_out/k3_blocks.txt
for mb in default 10 12 14 16; do
  if [ $mb = default ]; then unset DOKO_CUDA_LIB; else export DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_det$mb.so; fi
  echo "lib=$mb" >> gpurun_out/k3_blocks.txt
  timeout 300 python profiles/experiments/k3_sizes.py >> gpurun_out/k3_blocks.txt 2>&1
done
