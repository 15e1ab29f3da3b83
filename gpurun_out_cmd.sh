set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python profiles/bench_kernels.py > gpurun_out/kernels_v11.json 2> gpurun_out/kernels_v11.err; head -c 600 gpurun_out/kernels_v11.json; tail -3 gpurun_out/kernels_v11.err
python profiles/profile_playout.py > gpurun_out/plain_playout.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fdo_playout_fresh -s 2 -c 1 -o gpurun_out/prof_fdo_playout_v6 python profiles/profile_playout.py > gpurun_out/ncu_playout.log 2>&1
tail -2 gpurun_out/ncu_playout.log
