#!/usr/bin/env python3
"""Runs a few fdo playout launches (for ncu) and prints how many game steps one launch performs."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--engine", type=int, default=pkg.DK_FDO)
    ap.add_argument("--flags", type=int, default=1)
    a = ap.parse_args()
    dk = pkg.DokoCuda(0)
    tot = 0
    for i in range(a.launches):
        pts, steps = dk.playout(a.engine, a.n, dk.rng(0xD0C05EED, 0, 1000), flags=a.flags)
        torch.cuda.synchronize()
        tot = int(steps.sum(dtype=torch.int64))
    print(json.dumps({"n_games": a.n, "game_steps_per_launch": tot, "launches": a.launches, "engine": a.engine, "flags": a.flags}))


if __name__ == "__main__":
    main()
