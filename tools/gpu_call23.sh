set -x
mkdir -p gpurun_out
export DOKO_CUDA_UCT_PARTS=1
ncu --set full --clock-control none --import-source on -k regex:uct_tree -s 200 -c 1 -f -o gpurun_out/r02_uct_tree_v8 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_tree.log 2>&1; tail -2 gpurun_out/ncu_uct_tree.log
