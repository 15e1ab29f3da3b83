set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fdo_leaf_rollouts -s 1 -c 1 -f -o gpurun_out/r02_k4_roll_v1 python profiles/profile_kernels.py --which k4 > gpurun_out/ncu_k4.log 2>&1; tail -2 gpurun_out/ncu_k4.log
