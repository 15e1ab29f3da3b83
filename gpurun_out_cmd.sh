set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu_check.py > gpurun_out/r02_multigpu_check_2gpu_v13.txt 2>&1; tail -3 gpurun_out/r02_multigpu_check_2gpu_v13.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_v13_2gpu.json 2> gpurun_out/r02_bench_v13_2gpu.err; head -c 250 gpurun_out/r02_bench_v13_2gpu.json
