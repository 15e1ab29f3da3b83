set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err; tail -c 600 gpurun_out/bench_v14.json
python bench.py --impl reference > gpurun_out/bench_v14_reference_arm.json 2> gpurun_out/bench_v14_reference_arm.err; tail -c 400 gpurun_out/bench_v14_reference_arm.json
python profiles/experiments/state_ops_bw.py > gpurun_out/state_ops_bw.txt 2>&1; tail -1 gpurun_out/state_ops_bw.txt
python profiles/bench_kernels.py > gpurun_out/kernels_v39.json 2> gpurun_out/kernels_v39.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v14.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_v14.log 2>&1
