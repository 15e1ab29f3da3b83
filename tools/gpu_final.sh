# Round-end capture: GPU tests, bench lines (own arm + reference arm), ncu launch list of the bench, full ncu captures of the two
# kernels whose counters feed bench.py's roofline blocks (K2 at 2^24 games, K3 at 4096 x 4096 samples).
set -x
mkdir -p gpurun_out
V=${V:-v4}
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final_$V.log 2>&1; tail -3 gpurun_out/r02_pytest_final_$V.log
python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/r02_k2_plain_$V.json 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fdo_playout_fresh -s 1 -c 1 -f -o gpurun_out/r02_k2_$V python profiles/profile_playout.py --n 16777216 --launches 2 > gpurun_out/ncu_k2.log 2>&1
python profiles/experiments/det_profile_run.py > gpurun_out/r02_k3_plain_$V.json 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fdo_determinize -s 4 -c 1 -f -o gpurun_out/r02_k3_$V python profiles/experiments/det_profile_run.py > gpurun_out/ncu_k3.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_$V.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*$V*
