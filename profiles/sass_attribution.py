#!/usr/bin/env python3
"""Per-source-line and per-pipe attribution of one kernel's executed warp instructions.

Joins the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv`: executed count per SASS instruction) with the line
table of the cubin (`nvdisasm -g -c`), instruction by instruction (same order).  Pipes follow the measured rates in
profiles/r01_int_rates.json: ALU (IADD3 / LOP3 / SHF / SEL / ISETP / PRMT / VIMNMX ...) and FMA (IMAD*) issue one warp instruction
per two cycles per sub-partition each, IMAD.HI / IMAD.WIDE and LDS / SHFL one per four, POPC / FLO / BREV one per eight.

  python profiles/sass_attribution.py gpurun_out/prof.ncu-rep master_doko_reinforcement_learning_b200/libdoko_cuda.so <mangled-name> [steps]
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile


def pipe_of(op):
    base = op.split(".")[0]
    if base in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP", "IDP4A"):
        if op.startswith("IMAD.HI") or op.startswith("IMAD.WIDE"):
            return "fma_wide"
        return "fma"
    if base in ("POPC", "FLO", "BREV", "MUFU", "I2F", "F2I", "I2FP", "F2FP", "F2F"):
        return "xu"
    if base in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMS", "ATOMG", "RED", "SHFL", "LDSM", "LDC", "LDCU", "MATCH", "ATOM"):
        return "lsu"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "CALL", "RET", "BREAK", "NANOSLEEP", "BMOV", "JMP", "BRX", "JMX", "YIELD"):
        return "cbu"
    if base.startswith("U") and base not in ("UNPACK",) or base in ("S2UR", "R2UR", "VOTEU"):
        return "uniform"
    if base in ("NOP", "S2R", "CS2R", "DEPBAR", "ERRBAR", "MEMBAR"):
        return "misc"
    return "alu"


def main():
    rep, lib, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    steps = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = raw.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
    s = next(i for i, l in enumerate(dis) if l.startswith(kernel + ":"))
    insts, cur = [], ("?", 0)
    for l in dis[s + 1:]:
        if l.startswith("\t.section") or (l.startswith("//-----") and insts):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            txt = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            insts.append((cur, txt.split()[0]))
    assert len(insts) == len(rows), (len(insts), len(rows))
    by_line, by_pipe, by_op, by_file = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    line_pipe = collections.defaultdict(collections.Counter)
    total = thr = 0
    for (loc, op), r in zip(insts, rows):
        n = int(r["Instructions Executed"])
        total += n
        thr += int(r["Thread Instructions Executed"])
        by_line[loc] += n
        by_pipe[pipe_of(op)] += n
        by_op[op] += n
        by_file[loc[0]] += n
        line_pipe[loc][pipe_of(op)] += n
    out = {
        "kernel": kernel, "warp_inst": total, "thread_inst": thr, "avg_lanes": thr / total,
        "by_pipe_pct": {k: round(100.0 * v / total, 2) for k, v in by_pipe.most_common()},
        "pipe_cycles_per_issue_slot": {
            "alu": round(2.0 * by_pipe["alu"] / total, 3), "fma": round((2.0 * by_pipe["fma"] + 4.0 * by_pipe["fma_wide"]) / total, 3),
            "xu": round(8.0 * by_pipe["xu"] / total, 3), "lsu": round(4.0 * by_pipe["lsu"] / total, 3)},
        "top_ops_pct": {k: round(100.0 * v / total, 2) for k, v in by_op.most_common(24)},
        "top_lines_pct": [{"loc": f"{k[0]}:{k[1]}", "pct": round(100.0 * v / total, 2),
                           "pipes": {p: round(100.0 * c / total, 2) for p, c in line_pipe[k].most_common(3)}} for k, v in by_line.most_common(60)],
    }
    if steps:
        out["warp_inst_per_game_step"] = total / steps
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
