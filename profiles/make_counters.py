#!/usr/bin/env python3
"""Writes / updates profiles/kernel_counters.json from ncu_summary.py digests.

    python profiles/make_counters.py <mangled kernel name> <summary.json> <units per launch> <unit name> [source label] [games per launch]

Takes the LAST launch of the kernel in the digest (warmed up), records warp instructions per unit, pipe / issue utilisation, DRAM
bytes and the capture's duration, and ties them to the sha256 of the kernel's SASS in the library that was profiled
(master_doko_reinforcement_learning_b200/libdoko_cuda.sass.json, written by the build).  bench.py reports `roofline.stale: true` when the
loaded library's SASS differs from the one the counters were captured on."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
MS = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}


def scaled(k, key, table):
    v = k.get(key)
    if v is None or not isinstance(v, float):
        return None
    return v * table.get(k.get(key + "#unit", ""), 1.0)


def main():
    from master_doko_reinforcement_learning_b200 import _build

    mangled, summary, units, unit_name = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    label = sys.argv[5] if len(sys.argv) > 5 else summary
    dig = json.load(open(summary))
    ks = [k for k in dig["kernels"] if k.get("smsp__inst_executed.sum")]
    k = ks[-1]
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    data = json.load(open(path)) if os.path.exists(path) else {"kernels": {}}
    hashes = _build.sass_hashes() or _build.write_sass_hashes() or {}
    data["kernels"][mangled] = {
        "kernel": k["name"], "unit": unit_name, "capture_units": units,
        "warp_inst_per_unit": k["smsp__inst_executed.sum"] / units,
        "thread_inst_per_unit": k.get("smsp__thread_inst_executed.sum", 0.0) / units,
        "lanes_per_warp_inst": k.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "alu_pipe_pct_of_peak": k.get("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
        "issue_active_pct": k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_bytes_per_launch": ((scaled(k, "dram__bytes_read.sum", BYTES) or 0.0) + (scaled(k, "dram__bytes_write.sum", BYTES) or 0.0)) or None,
        "capture_kernel_ms": scaled(k, "gpu__time_duration.sum", MS),
        "capture_games": float(sys.argv[6]) if len(sys.argv) > 6 else (units if unit_name == "game" else None),
        "registers_per_thread": k.get("launch__registers_per_thread"),
        "source": label, "sass_sha256": hashes.get(mangled),
    }
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data["kernels"][mangled], indent=1))


if __name__ == "__main__":
    main()
