set -x
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_v12_${N}gpu.json 2> gpurun_out/bench_v12_${N}gpu.err; tail -1 gpurun_out/bench_v12_${N}gpu.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
nvidia-smi topo -m 2>/dev/null | head -14
cat /sys/devices/system/node/node*/cpulist
