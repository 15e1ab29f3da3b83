set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final_v15.log 2>&1; tail -3 gpurun_out/r02_pytest_final_v15.log
python -c "import __graft_entry__ as g; g.smoke()"
