V=v14 bash tools/gpu_final.sh
timeout 200 python profiles/bench_kernels.py > gpurun_out/r02_kernels_v18.json 2> gpurun_out/r02_kernels_v18.err; head -c 1500 gpurun_out/r02_kernels_v18.json
