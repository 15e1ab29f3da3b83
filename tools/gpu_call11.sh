set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_state_ops.py tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_v11.log 2>&1; tail -5 gpurun_out/r02_pytest_v11.log
for i in 1 2; do python profiles/experiments/state_ops_bw.py 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('t2', d['apply'])"; done
for i in 1 2; do DOKO_CUDA_LIB=$PWD/build/variants/libdoko_cuda_apply_t1.so python profiles/experiments/state_ops_bw.py 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('t1', d['apply'])"; done
