set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_v1.log 2>&1; tail -15 gpurun_out/r02_pytest_v1.log
python bench.py > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; tail -c 600 gpurun_out/r02_bench_v1.err; tail -c 300 gpurun_out/r02_bench_v1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_v1_reference.json 2> gpurun_out/r02_bench_v1_reference.err
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/r02_k2_plain.json 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fdo_playout_fresh -s 1 -c 1 -o gpurun_out/r02_k2_v1 python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/ncu_k2.log 2>&1
python profiles/experiments/det_profile_run.py > gpurun_out/r02_k3_plain.json 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fdo_determinize -s 4 -c 1 -o gpurun_out/r02_k3_v1 python profiles/experiments/det_profile_run.py > gpurun_out/ncu_k3.log 2>&1
ls -la gpurun_out | tail -20
