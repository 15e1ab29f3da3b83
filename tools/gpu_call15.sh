set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_pimc.py -m gpu -x -q > gpurun_out/r02_pytest_v13.log 2>&1; tail -3 gpurun_out/r02_pytest_v13.log
for p in 1 2 3 4; do
DOKO_CUDA_UCT_PARTS=$p python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v7_parts$p.json 2> gpurun_out/r02_uct_bench_v7.err; python -c "
import json; d=json.load(open('gpurun_out/r02_uct_bench_v7_parts$p.json')); print('parts $p', {k: '%.3g'%v['iterations_per_s'] for k,v in d.items()})"
done
