set -x
mkdir -p gpurun_out
python profiles/profile_kernels.py --which uct > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 120 --csv --log-file gpurun_out/r02_uct_launches_v6.csv python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct.log 2>&1
tail -2 gpurun_out/ncu_uct.log
ncu --set full --clock-control none --import-source on -k regex:uct_tree -s 200 -c 1 -f -o gpurun_out/r02_uct_tree_v6 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_tree.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uct_rollout -s 200 -c 1 -f -o gpurun_out/r02_uct_rollout_v6 python profiles/profile_kernels.py --which uct > gpurun_out/ncu_uct_roll.log 2>&1
ls -la gpurun_out/*v6*
