set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_v10.log 2>&1; tail -5 gpurun_out/r02_pytest_v10.log
python profiles/experiments/state_ops_bw.py > gpurun_out/r02_state_ops_bw_v2.json 2>&1; cat gpurun_out/r02_state_ops_bw_v2.json
python profiles/bench_kernels.py > gpurun_out/r02_kernels_v3.json 2> gpurun_out/r02_kernels_v3.err; tail -c 2500 gpurun_out/r02_kernels_v3.json
python profiles/experiments/state_ops_bw.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:apply_tma -s 3 -c 1 -o gpurun_out/r02_apply_v2 python profiles/experiments/state_ops_bw.py > gpurun_out/ncu_apply.log 2>&1
