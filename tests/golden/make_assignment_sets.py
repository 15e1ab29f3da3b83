#!/usr/bin/env python3
"""Extract the support-set tests of rs-doko-assignment (rs-doko-assignment/src/assignment.rs:908-1186) into
tests/golden/assignment_sets.json: inputs of `brute_force_assignments(...)` and the enumerated set of assignments that the 1000
samples must equal (RNG independent).  Run in the build container only."""
import json
import os
import re

REF = "/root/reference/rs-doko-assignment/src/assignment.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assignment_sets.json")
CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
PL = {"PLAYER_BOTTOM": 0, "PLAYER_LEFT": 1, "PLAYER_TOP": 2, "PLAYER_RIGHT": 3}


def main():
    src = re.sub(r"//[^\n]*", "", open(REF, encoding="utf-8").read())
    src = src[src.index("fn brute_force_assignments("):]
    cases = []
    for m in re.finditer(r"fn (test_\w+)\(\) \{(.*?)\n    \}\n", src, re.S):
        name, body = m.group(1), m.group(2)
        call = re.search(r"brute_force_assignments\(\s*(None|Some\((\w+)\)),\s*&\[(.*?)\],\s*hand_from_vec\(vec!\[(.*?)\]\),\s*&\[(\d+), (\d+), (\d+), (\d+)\],\s*(\w+),", body, re.S)
        if not call:
            continue
        tricks = [{"start": PL[a], "cards": [CARD_ID[c] for c in re.findall(r"DoCard::(\w+)", b)]}
                  for a, b in re.findall(r"DoTrick::existing\((\w+), vec!\[(.*?)\]\)", call.group(3), re.S)]
        n = re.search(r"assert_eq!\(result\.len\(\), (\d+)\)", body)
        sets = []
        for am in re.finditer(r"result\.contains\(&\[(.*?)\]\)\);", body, re.S):
            hands = [[CARD_ID[c.strip()] for c in h.split(",") if c.strip()] for h in re.findall(r"vec!\[(.*?)\]", am.group(1))]
            sets.append([sorted(h) for h in hands])
        cases.append({"name": name, "marriage": PL[call.group(2)] if call.group(2) else -1, "tricks": tricks,
                      "hand": [CARD_ID[c] for c in re.findall(r"DoCard::(\w+)", call.group(4))], "lens": [int(call.group(i)) for i in range(5, 9)],
                      "observer": PL[call.group(9)], "n_unique": int(n.group(1)) if n else None, "assignments": sets})
    json.dump({"source": "rs-doko-assignment/src/assignment.rs:908-1186", "cases": cases}, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, [(c["name"], c["n_unique"], len(c["assignments"])) for c in cases])


if __name__ == "__main__":
    main()
