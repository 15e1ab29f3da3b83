set -x
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err ) 2> gpurun_out/bench_v18.time; tail -3 gpurun_out/bench_v18.time; tail -c 400 gpurun_out/bench_v18.err
