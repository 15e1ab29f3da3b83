set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_v8_8gpu.json 2> gpurun_out/r02_bench_v8_8gpu.err; tail -c 300 gpurun_out/r02_bench_v8_8gpu.err; head -c 300 gpurun_out/r02_bench_v8_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_bench_v8_8gpu_reference.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_v8_8gpu_reference.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_v8_2gpu.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_v8_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_v8_4gpu.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_v8_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu_check.py > gpurun_out/r02_multigpu_check_8gpu.txt 2>&1; tail -2 gpurun_out/r02_multigpu_check_8gpu.txt
python bench.py > gpurun_out/r02_bench_v8_1gpu.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_v8_1gpu.json
