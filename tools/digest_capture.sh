#!/bin/bash
# tools/digest_capture.sh <V>  — after tools/gpu_final.sh (V=<V>) has run on the GPU box: turns gpurun_out/r02_{k2,k3}_<V>.ncu-rep and the
# bench's launch list into the tracked summaries under profiles/ and refreshes profiles/kernel_counters.json (tied to the SASS hashes of
# the library in the tree, which must be the one that was profiled).
set -e
V=$1
cd "$(dirname "$0")/.."
steps=$(python -c "import json;print(json.load(open('gpurun_out/r02_k2_plain_$V.json'))['game_steps_per_launch'])")
python profiles/ncu_summary.py gpurun_out/r02_k2_$V.ncu-rep $steps > profiles/r02_k2_${V}_ncu_summary.json
python profiles/ncu_summary.py gpurun_out/r02_k3_$V.ncu-rep > profiles/r02_k3_${V}_ncu_summary.json
python profiles/make_counters.py _ZN2dk24fdo_playout_fresh_kernelILb1EEEvNS_9RngParamsEmPvS2_jPy profiles/r02_k2_${V}_ncu_summary.json $steps "game step" \
  "profiles/r02_k2_${V}_ncu_summary.json (ncu --set full --clock-control none, 2^24 games per launch)" 16777216 > /dev/null
python profiles/make_counters.py _ZN2dk22fdo_determinize_kernelENS_9RngParamsEmjjPK8dk_statePmPhS5_ profiles/r02_k3_${V}_ncu_summary.json 16777216 sample \
  "profiles/r02_k3_${V}_ncu_summary.json (ncu --set full, 4096 info-states x 4096 samples per launch)" > /dev/null
cp gpurun_out/r02_launches_bench_$V.csv profiles/r02_launches_bench_$V.csv
python - "$V" <<'PY'
import csv, collections, sys
V = sys.argv[1]
rows = [r for r in csv.reader(open(f"profiles/r02_launches_bench_{V}.csv")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[iv] in ("", "Metric Value"):
        continue
    try:
        ms = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    except ValueError:
        continue
    name = r[ik].split("(")[0][:260]
    tot[name] += ms; cnt[name] += 1
allms = sum(tot.values())
with open(f"profiles/r02_launches_bench_{V}_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 1 (first 600 launches: the K2 warm-up + timed steps, the end-to-end forms, config 0, the determinization passes ...)\n")
    f.write("kernel | launches | total ms | share of the captured launches\n")
    for k, v in tot.most_common():
        f.write(f"{k} | {cnt[k]} | {v:.3f} | {100 * v / allms:.1f} %\n")
    f.write("(the timed region of `value` launches fdo_playout_fresh_kernel<1> only: gpu_launches = steps; every other kernel belongs to an extra key of the line; the torch index / gather kernels prepare the mid-game states of the extra configs outside every timed region)\n")
PY
head -6 profiles/r02_launches_bench_${V}_summary.txt
