set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -c 3000 gpurun_out/bench_v6.json; tail -3 gpurun_out/bench_v6.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v6_ref.json 2> gpurun_out/bench_v6_ref.err; tail -c 1200 gpurun_out/bench_v6_ref.json
python bench.py --steps 2 --warmup 1 > gpurun_out/b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
