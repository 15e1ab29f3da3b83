set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python profiles/bench_kernels.py > gpurun_out/kernels3.json 2> gpurun_out/kernels3.err; tail -2 gpurun_out/kernels3.err; cat gpurun_out/kernels3.json
