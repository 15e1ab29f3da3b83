set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matching.py tests/test_gpu_pimc.py tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_v15.log 2>&1; tail -3 gpurun_out/r02_pytest_v15.log
python profiles/experiments/det_profile_run.py > gpurun_out/r02_k3_plain_v2.json 2>&1; cat gpurun_out/r02_k3_plain_v2.json
ncu --set full --clock-control none --import-source on -k regex:fdo_determinize -s 2 -c 1 -f -o gpurun_out/r02_k3_v2 python profiles/experiments/det_profile_run.py > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
