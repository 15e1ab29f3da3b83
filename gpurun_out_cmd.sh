set -x
mkdir -p gpurun_out
python profiles/experiments/enc_doko.py > gpurun_out/enc_doko.txt 2>&1; cat gpurun_out/enc_doko.txt
