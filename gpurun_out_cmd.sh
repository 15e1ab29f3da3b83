set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v12_1gpu.json 2> gpurun_out/bench_v12_1gpu.err; tail -1 gpurun_out/bench_v12_1gpu.json | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v12_reference_arm.json 2>&1; tail -1 gpurun_out/bench_v12_reference_arm.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v12.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_v12.log 2>&1
