#!/usr/bin/env python3
"""Small single-purpose launches of K3 / K4 / K5 for ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 0xD0C05EED


def main():
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="k5")
    a = ap.parse_args()
    dk = pkg.DokoCuda(0)
    n = 1 << 20
    states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
    for k in range(30):
        dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
    torch.cuda.synchronize()
    if a.which == "k5":
        obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
        for k in range(3):
            dk.step_random_encode(states, dk.rng(SEED, 0, 100 + k), obs_out=obs)
        dk.encode(pkg.DK_LAYOUT_FDO_PI311, states, out=obs)
    elif a.which == "k3":
        sub = states[:1024].clone()
        for k in range(3):
            dk.determinize(pkg.DK_FDO, sub, 4096, dk.rng(SEED, 0, 9))
    elif a.which == "k4":
        sub = states[:1024].clone()
        for k in range(3):
            dk.leaf_rollouts(sub, 1024, dk.rng(SEED, 0, 11), determinize=True)
    torch.cuda.synchronize()
    print("done", a.which)


if __name__ == "__main__":
    main()
