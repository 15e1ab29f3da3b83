set -x
mkdir -p gpurun_out
python profiles/experiments/assign_profile_run.py > gpurun_out/r02_assign_plain.json 2>&1; cat gpurun_out/r02_assign_plain.json
ncu --set full --clock-control none --import-source on -k regex:doko_assign -s 2 -c 1 -f -o gpurun_out/r02_assign_v1 python profiles/experiments/assign_profile_run.py > gpurun_out/ncu_assign.log 2>&1; tail -2 gpurun_out/ncu_assign.log
