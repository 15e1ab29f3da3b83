#!/bin/bash
# tools/build_variant.sh <name> [-DFLAG=VALUE ...]  →  build/variants/libdoko_cuda_<name>.so (an experimental build of the same ABI;
# select it with DOKO_CUDA_LIB=build/variants/libdoko_cuda_<name>.so).  Prints the kernels' register / spill lines that match $GREP.
set -e
name=$1; shift
mkdir -p build/variants
cd master_doko_reinforcement_learning_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -Xptxas -v "$@" \
  -o ../../build/variants/libdoko_cuda_$name.so cabi.cu -ldl 2>&1 | grep -A2 -E "${GREP:-uct_}" | grep -E "Compiling|Used|spill" | sed 's/ptxas info    : //'
