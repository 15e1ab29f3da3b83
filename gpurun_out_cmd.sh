set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_v11_8gpu.json 2> gpurun_out/r02_bench_v11_8gpu.err; tail -c 300 gpurun_out/r02_bench_v11_8gpu.err; head -c 300 gpurun_out/r02_bench_v11_8gpu.json
