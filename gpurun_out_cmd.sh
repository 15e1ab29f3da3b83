set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v20.json 2> gpurun_out/bench_v20.err; tail -c 200 gpurun_out/bench_v20.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v20_reference_arm.json 2> gpurun_out/bench_v20_reference_arm.err; tail -c 300 gpurun_out/bench_v20_reference_arm.json
