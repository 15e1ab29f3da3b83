set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -2 gpurun_out/bench3.err; cat gpurun_out/bench3.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench3_ref.json 2>&1; cat gpurun_out/bench3_ref.json
python profiles/bench_kernels.py > gpurun_out/kernels2.json 2> gpurun_out/kernels2.err; cat gpurun_out/kernels2.json
python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:fdo_playout -s 1 -c 1 -o gpurun_out/prof_fdo_playout_v3 -f python profiles/profile_playout.py --n 1048576 --launches 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
