set -x
mkdir -p gpurun_out
python profiles/profile_kernels.py --which uct > gpurun_out/plain_uct.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:uct_select -s 200 -c 1 -o gpurun_out/r02_uct_select_v1 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_sel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_expand -s 200 -c 1 -o gpurun_out/r02_uct_expand_v1 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_exp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_rollout -s 200 -c 1 -o gpurun_out/r02_uct_rollout_v1 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_roll.log 2>&1
ls -la gpurun_out/*.ncu-rep
