set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'apply_tma' --launch-skip 3 -c 1 -o gpurun_out/prof_apply_tma -f python profiles/experiments/state_ops_bw.py > gpurun_out/ncu_apply.log 2>&1; tail -2 gpurun_out/ncu_apply.log
