set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'legal_mask_tma|apply_tma' --launch-skip 4 -c 2 -o gpurun_out/prof_stateops_tma -f python profiles/experiments/state_ops_bw.py > gpurun_out/ncu_stateops.log 2>&1; tail -2 gpurun_out/ncu_stateops.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pack_replay --launch-skip 2 -c 1 -o gpurun_out/prof_packer_v3 -f python profiles/experiments/n4_bench.py > gpurun_out/ncu_packer3.log 2>&1; tail -2 gpurun_out/ncu_packer3.log
