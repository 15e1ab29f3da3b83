set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err; cat gpurun_out/bench5.json | cut -c1-1600
python bench.py --steps 2 --warmup 1 > gpurun_out/bench5_short.json 2>&1; cut -c1-400 gpurun_out/bench5_short.json
python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:fdo_playout -s 1 -c 1 -o gpurun_out/prof_fdo_playout_v4 -f python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
