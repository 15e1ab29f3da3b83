"""Extracts the reference's PolicyFusion known-answer tests into tests/golden/policy_fusion_cases.json.

Source: /root/reference/rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:125-313 (the four #[tokio::test] functions) and the
FdoAction::to_index table (/root/reference/rs-full-doko/src/action/action.rs:56-107).  Run in the build container only; the
GPU box reads the committed JSON.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def action_indices():
    src = open(f"{REF}/rs-full-doko/src/action/action.rs").read()
    body = src[src.index("pub fn to_index"):src.index("pub fn from_index")]
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"FdoAction::(\w+)\s*=>\s*(\d+)", body)}


def main():
    idx = action_indices()
    assert len(idx) == 39
    src = open(f"{REF}/rs-doko-py-bridge/src/compare_impi/policy_fusion.rs").read()
    tests_src = src[src.index("#[cfg(test)]"):]
    cases = []
    for m in re.finditer(r"async fn (test_\w+)\(\)\s*\{(.*?)\n    \}\n", tests_src, re.S):
        name, body = m.group(1), m.group(2)
        strategy = re.search(r"let fusion = (\w+)\s*\{", body).group(1)
        var = {v: idx[a] for v, a in re.findall(r"let (\w+)\s*=\s*FdoAction::(\w+)\.to_index\(\)", body)}

        def resolve(tok):
            return var[tok] if tok in var else int(tok)

        rows_src = body[body.index("vec!["):body.index("let dummy_values")]
        rows = []
        for blk in re.finditer(r"\{\s*let mut s = dummy;(.*?)\bs\s*\}", rows_src, re.S):
            row = [0] * 39
            for k, v in re.findall(r"s\[(\w+)\]\s*=\s*(\d+)\s*;", blk.group(1)):
                row[resolve(k)] = int(v)
            rows.append(row)
        allowed_src = body[body.index("from_vec("):body.index("await")]
        allowed = [idx[a] for a in re.findall(r"FdoAction::(Card\w+|Reservation\w+|Announcement\w+|NoAnnouncement)\b", allowed_src)]
        allowed += [int(i) for i in re.findall(r"FdoAction::from_index\((\d+)\)", allowed_src)]
        exp_src = body[body.index("assert_eq!("):]
        e = re.search(r"result,\s*FdoAction::from_index\((\d+)\)", exp_src)
        expected = int(e.group(1)) if e else idx[re.search(r"result,\s*FdoAction::(\w+)", exp_src).group(1)]
        cases.append({"name": name, "strategy": strategy, "visits": rows, "allowed": sorted(allowed), "expected": expected})
    assert len(cases) == 4, len(cases)
    with open(os.path.join(HERE, "policy_fusion_cases.json"), "w") as f:
        json.dump({"source": "rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:125-313", "cases": cases}, f, indent=1)
    for c in cases:
        print(c["name"], c["strategy"], len(c["visits"]), c["allowed"], "->", c["expected"])


if __name__ == "__main__":
    main()
