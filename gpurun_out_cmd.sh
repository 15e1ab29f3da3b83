set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -2 gpurun_out/bench2.err; cat gpurun_out/bench2.json
python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:fdo_playout -s 1 -c 1 -o gpurun_out/prof_fdo_playout_v2 -f python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
