set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python profiles/bench_kernels.py > gpurun_out/kernels1.json 2> gpurun_out/kernels1.err; tail -3 gpurun_out/kernels1.err; cat gpurun_out/kernels1.json
