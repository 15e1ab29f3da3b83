set -x
mkdir -p gpurun_out
timeout 300 python tests/soak_parity.py --seconds 40 > gpurun_out/soak_parity2.json 2> gpurun_out/soak_parity2.err; tail -c 600 gpurun_out/soak_parity2.json; tail -c 300 gpurun_out/soak_parity2.err
