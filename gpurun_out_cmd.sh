set -x
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:fdo_playout_fresh -s 1 -c 1 -f -o gpurun_out/prof_k2_v10_2p24 python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/prof_k2_v10.log 2>&1
python profiles/bench_kernels.py > gpurun_out/kernels_v29.json 2> gpurun_out/kernels_v29.err
