set -x
mkdir -p gpurun_out
for c in "" 100 75 66 50 40 25; do if [ -z "$c" ]; then python profiles/experiments/k5_carveout.py; else DK_ENC_CARVEOUT=$c python profiles/experiments/k5_carveout.py; fi; done > gpurun_out/carveout.txt 2>&1
cat gpurun_out/carveout.txt
