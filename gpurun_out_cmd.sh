set -x
mkdir -p gpurun_out
python profiles/bench_kernels.py > gpurun_out/kernels_v6.json 2> gpurun_out/kernels_v6.err; tail -c 900 gpurun_out/kernels_v6.json; tail -3 gpurun_out/kernels_v6.err
