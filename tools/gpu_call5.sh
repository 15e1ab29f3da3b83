set -x
mkdir -p gpurun_out
export DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_uct_b.so
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:uct_select -s 200 -c 1 -o gpurun_out/r02_uct_select_v2 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_sel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_expand -s 200 -c 1 -o gpurun_out/r02_uct_expand_v2 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_exp.log 2>&1
