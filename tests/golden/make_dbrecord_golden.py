#!/usr/bin/env python3
"""Writes tests/golden/dbrecord_golden.json: three experience rows (the reference's own 311-token encode_state_pi vector among them)
and their DBRecord bytes as produced by the independent bincode encoder tests/golden/bincode_v1.py."""
import base64
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import sys

    sys.path.insert(0, HERE)
    import bincode_v1

    pi = json.load(open(os.path.join(HERE, "encode_pi_vector.json")))
    tokens = pi["expected"] if isinstance(pi, dict) and "expected" in pi else None
    prng = np.random.default_rng(20261019)
    rows = []
    for k in range(3):
        state = [int(x) for x in (tokens if (k == 0 and tokens and len(tokens) == 311) else prng.integers(-3, 64, size=311))]
        value = [float(np.float32(x)) for x in (prng.integers(-96, 97, size=4) / 8.0)]
        policy = prng.random(39).astype(np.float32)
        policy = [float(x) for x in (policy / policy.sum()).astype(np.float32)]
        rec = bincode_v1.db_record(state, value, policy)
        assert len(rec) == 2684
        rows.append({"state": state, "value": value, "policy": policy, "bincode_b64": base64.b64encode(rec).decode()})
    json.dump({"format": "bincode 1.3.3 DefaultOptions (fixint, little endian), heapless 0.8.0 Vec as seq", "rows": rows},
              open(os.path.join(HERE, "dbrecord_golden.json"), "w"))
    print("wrote dbrecord_golden.json")


if __name__ == "__main__":
    main()
