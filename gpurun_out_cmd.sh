set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python profiles/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; tail -1 gpurun_out/sanitize_plain.log
python bench.py > gpurun_out/bench_v17.json 2> gpurun_out/bench_v17.err; tail -c 300 gpurun_out/bench_v17.json
python bench.py --impl reference > gpurun_out/bench_v17_reference_arm.json 2> gpurun_out/bench_v17_reference_arm.err; tail -c 200 gpurun_out/bench_v17_reference_arm.json
python profiles/bench_kernels.py > gpurun_out/kernels_v41.json 2> gpurun_out/kernels_v41.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v17.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_v17.log 2>&1
