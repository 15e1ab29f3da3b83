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
    elif a.which == "uct":
        sub = states[:2048].clone()
        need = dk.L.dk_uct_workspace_bytes(2048 * 64, 256)
        ws = torch.empty((need // 8 + 2,), dtype=torch.int64, device="cuda")
        for k in range(2):
            dk.uct_search(sub, 256, 1.4, dk.rng(SEED, 0, 13), trees_per_root=64, determinize=True, workspace=ws)
    elif a.which == "pimc":
        sub = states[:1024].clone()
        for k in range(3):
            dk.pimc_evaluate(sub, 64, 32, dk.rng(SEED, 0, 12), want_values=False)
    elif a.which == "sp":
        sp = dk.self_play(n, n * 3)
        for k in range(3):
            sp.reset()
            sp.begin_turn(states, 0, 1.0, dk.rng(SEED, 0, 200 + k))
            sp.uniform_search(dk.rng(SEED, 0, 200 + k))
            sp.end_turn(states)
    torch.cuda.synchronize()
    print("done", a.which)


if __name__ == "__main__":
    main()
