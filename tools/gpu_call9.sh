set -x
mkdir -p gpurun_out
python profiles/experiments/state_ops_bw.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:apply_tma -s 3 -c 1 -o gpurun_out/r02_apply_sorted python profiles/experiments/state_ops_bw.py > gpurun_out/ncu_apply.log 2>&1
tail -3 gpurun_out/ncu_apply.log
