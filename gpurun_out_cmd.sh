set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fdo_step_encode_tma --launch-skip 33 -c 1 -o gpurun_out/prof_k5_tma -f python profiles/experiments/k5_tma.py > gpurun_out/ncu_k5.log 2>&1; tail -2 gpurun_out/ncu_k5.log
