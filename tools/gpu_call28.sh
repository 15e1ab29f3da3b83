set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_v4_8gpu.json 2> gpurun_out/r02_bench_v4_8gpu.err; tail -c 400 gpurun_out/r02_bench_v4_8gpu.err; head -c 600 gpurun_out/r02_bench_v4_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tests/multigpu_check.py > gpurun_out/r02_multigpu_check_8gpu.txt 2>&1; tail -3 gpurun_out/r02_multigpu_check_8gpu.txt
