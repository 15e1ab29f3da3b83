set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python profiles/bench_kernels.py > gpurun_out/kernels_v19.json 2> gpurun_out/kernels_v19.err; head -c 900 gpurun_out/kernels_v19.json
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v8_1gpu.json 2> gpurun_out/bench_v8_1gpu.err; tail -1 gpurun_out/bench_v8_1gpu.json
python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/prof_k2_v8.steps.json 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:fdo_playout_fresh -s 1 -c 1 -f -o gpurun_out/prof_k2_v8_2p24 python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/prof_k2_v8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v8.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_v8.log 2>&1
