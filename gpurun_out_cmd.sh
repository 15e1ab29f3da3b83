set -x
mkdir -p gpurun_out
timeout 300 python tests/soak_parity.py --seconds 110 --samplers > gpurun_out/r02_soak_parity_v11.json 2> gpurun_out/r02_soak_v11.err; tail -c 700 gpurun_out/r02_soak_parity_v11.json; tail -3 gpurun_out/r02_soak_v11.err
python bench.py > gpurun_out/r02_bench_v11_1gpu.json 2> gpurun_out/r02_bench_v11_1gpu.err; tail -c 300 gpurun_out/r02_bench_v11_1gpu.err; head -c 600 gpurun_out/r02_bench_v11_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_v11_reference.json 2> gpurun_out/r02_bench_v11_reference.err; head -c 300 gpurun_out/r02_bench_v11_reference.json
timeout 200 python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v11.json 2> gpurun_out/r02_uct_bench_v11.err; head -c 900 gpurun_out/r02_uct_bench_v11.json
