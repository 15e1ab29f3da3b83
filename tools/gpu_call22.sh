set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_uct.py tests/test_gpu_pimc.py -m gpu -x -q > gpurun_out/r02_pytest_v19.log 2>&1; tail -3 gpurun_out/r02_pytest_v19.log
for v in main uct8_b7 uct8_b6; do
for p in 1 3; do
L=""; if [ $v != main ]; then L=$PWD/build/variants/libdoko_cuda_$v.so; fi
DOKO_CUDA_LIB=$L DOKO_CUDA_UCT_PARTS=$p python profiles/bench_uct.py > gpurun_out/r02_uct_bench_v8_${v}_parts$p.json 2> gpurun_out/r02_uct_bench_v8.err; python -c "
import json; d=json.load(open('gpurun_out/r02_uct_bench_v8_${v}_parts$p.json')); print('$v parts $p', {k: '%.3g'%v['iterations_per_s'] for k,v in d.items()})"
done; done
tail -3 gpurun_out/r02_uct_bench_v8.err
