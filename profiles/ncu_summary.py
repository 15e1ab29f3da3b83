#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into a small text/JSON digest for profiles/."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    game_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    out = {"report": rep, "kernels": []}
    for vals in rows[2:]:
        k = {"name": vals[hdr.index("Kernel Name")]}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                try:
                    k[key] = float(vals[i].replace(",", ""))
                except ValueError:
                    k[key] = vals[i]
                k[key + "#unit"] = units[i]
        if game_steps:
            k["game_steps_per_launch"] = game_steps
            k["warp_inst_per_game_step"] = k["smsp__inst_executed.sum"] / game_steps
            if "smsp__thread_inst_executed_per_inst_executed.ratio" in k:
                k["thread_inst_per_game_step"] = k["warp_inst_per_game_step"] * k["smsp__thread_inst_executed_per_inst_executed.ratio"]
        out["kernels"].append(k)
    # stall reasons
    for vals, k in zip(rows[2:], out["kernels"]):
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(vals[i])
                except ValueError:
                    pass
        k["stall_warps_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
