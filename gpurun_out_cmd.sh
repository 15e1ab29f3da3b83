set -x
mkdir -p gpurun_out
python profiles/profile_playout.py --n 16777216 --launches 3 > gpurun_out/plain_playout_2p24.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fdo_playout_fresh -s 2 -c 1 -o gpurun_out/prof_fdo_playout_v6_2p24 python profiles/profile_playout.py --n 16777216 --launches 3 > gpurun_out/ncu_playout_2p24.log 2>&1
tail -2 gpurun_out/ncu_playout_2p24.log; tail -1 gpurun_out/plain_playout_2p24.log
