set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v10_1gpu.json 2> gpurun_out/bench_v10_1gpu.err; tail -1 gpurun_out/bench_v10_1gpu.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['e2e']['int32_api']['value'], d['roofline']['frac'], d['roofline']['alu_pipe'])
"
