set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_v2_1gpu.json 2> gpurun_out/r02_bench_v2_1gpu.err; tail -c 300 gpurun_out/r02_bench_v2_1gpu.err; head -c 400 gpurun_out/r02_bench_v2_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_v2_2gpu.json 2> gpurun_out/r02_bench_v2_2gpu.err; tail -c 600 gpurun_out/r02_bench_v2_2gpu.err; head -c 300 gpurun_out/r02_bench_v2_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_v2_2gpu_reference.json 2> gpurun_out/r02_bench_v2_2gpu_reference.err; head -c 300 gpurun_out/r02_bench_v2_2gpu_reference.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/multigpu_check.py > gpurun_out/r02_multigpu_check_2gpu.txt 2>&1; tail -5 gpurun_out/r02_multigpu_check_2gpu.txt
