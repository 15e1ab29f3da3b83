set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_v22.log 2>&1; tail -2 gpurun_out/r02_pytest_v22.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_v5_reference.json 2> gpurun_out/r02_bench_v5_reference.err; head -c 300 gpurun_out/r02_bench_v5_reference.json; echo
python bench.py > gpurun_out/r02_bench_v5_1gpu.json 2> gpurun_out/r02_bench_v5_1gpu.err; tail -c 200 gpurun_out/r02_bench_v5_1gpu.err; head -c 300 gpurun_out/r02_bench_v5_1gpu.json; echo
