set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final_v13.log 2>&1; tail -3 gpurun_out/r02_pytest_final_v13.log
timeout 200 python profiles/experiments/narrow_rows.py > gpurun_out/r02_narrow_rows_final.json 2> gpurun_out/r02_narrow_rows.err; cat gpurun_out/r02_narrow_rows_final.json
python bench.py > gpurun_out/r02_bench_v13_1gpu.json 2> gpurun_out/r02_bench_v13_1gpu.err; tail -c 300 gpurun_out/r02_bench_v13_1gpu.err; head -c 400 gpurun_out/r02_bench_v13_1gpu.json
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v13.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
