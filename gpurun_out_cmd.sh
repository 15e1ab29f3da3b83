set -x
mkdir -p gpurun_out
for t in 128 256 512; do DOKO_CUDA_LIB=$PWD/profiles/experiments/libs/libdoko_pimc$t.so python profiles/experiments/pimc_cfg.py; done > gpurun_out/pimc_cfg.txt 2>&1; cat gpurun_out/pimc_cfg.txt
