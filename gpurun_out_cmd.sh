set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python profiles/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; tail -1 gpurun_out/sanitize_plain.log
python profiles/bench_kernels.py > gpurun_out/kernels_v43.json 2> gpurun_out/kernels_v43.err; tail -c 200 gpurun_out/kernels_v43.err
