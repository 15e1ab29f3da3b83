set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "narrow or lockstep or step_random" > gpurun_out/r02_pytest_v15.log 2>&1; tail -3 gpurun_out/r02_pytest_v15.log
DOKO_CUDA_NO_TMA=1 timeout 300 python -m pytest tests -m gpu -x -q -k "narrow" > gpurun_out/r02_pytest_v15_notma.log 2>&1; tail -2 gpurun_out/r02_pytest_v15_notma.log
timeout 200 python profiles/experiments/narrow_rows.py > gpurun_out/r02_narrow_rows_tma.json 2> gpurun_out/r02_narrow_rows.err; cat gpurun_out/r02_narrow_rows_tma.json
