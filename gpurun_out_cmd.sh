set -x
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/multigpu_check.py > gpurun_out/multigpu_check_${N}gpu.txt 2>&1; tail -2 gpurun_out/multigpu_check_${N}gpu.txt
timeout 300 $TR profiles/bench_scaling.py > gpurun_out/scaling_${N}gpu.json 2> gpurun_out/scaling_${N}gpu.err; tail -1 gpurun_out/scaling_${N}gpu.json
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-250
timeout 400 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_reference.json 2> gpurun_out/bench_${N}gpu_reference.err; tail -1 gpurun_out/bench_${N}gpu_reference.json | cut -c1-200
