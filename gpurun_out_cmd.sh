set -x
mkdir -p gpurun_out
python profiles/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; tail -3 gpurun_out/sanitize_plain.log
./profiles/experiments/oracle_inst_count oracle/liboracle.so > gpurun_out/oracle_inst_count.json 2>&1; cat gpurun_out/oracle_inst_count.json
cat /proc/sys/kernel/perf_event_paranoid
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v9_1gpu.json 2> gpurun_out/bench_v9_1gpu.err; tail -1 gpurun_out/bench_v9_1gpu.json | cut -c1-600
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v9_reference_arm.json 2>&1; tail -1 gpurun_out/bench_v9_reference_arm.json | cut -c1-300
