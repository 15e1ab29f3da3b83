set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_v14_1gpu.json 2> gpurun_out/r02_bench_v14_1gpu.err; tail -c 200 gpurun_out/r02_bench_v14_1gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_v14_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stale'], d['roofline']['frac'], d['determinizations']['roofline']['stale'], d['config3_leaf_rollouts']['value'], d['config3_leaf_rollouts']['determinize_every_rollout']['value'])"
python -c "import __graft_entry__ as g; g.smoke()"
