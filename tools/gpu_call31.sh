set -x
mkdir -p gpurun_out
for v in main t6r3 t8r3 t6r4 t5r3; do
L=""; if [ $v != main ]; then L=$PWD/build/variants/libdoko_cuda_$v.so; fi
DOKO_CUDA_LIB=$L python profiles/bench_uct.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', {k: '%.3g'%v['iterations_per_s'] for k,v in d.items()})"
done
