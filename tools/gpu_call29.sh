set -x
mkdir -p gpurun_out
timeout 400 python tests/soak_parity.py --seconds 150 --samplers > gpurun_out/r02_soak_parity.json 2> gpurun_out/r02_soak_parity.err; echo rc=$?; cat gpurun_out/r02_soak_parity.json; tail -5 gpurun_out/r02_soak_parity.err
