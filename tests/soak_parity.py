#!/usr/bin/env python3
"""Parity soak (B200): the fresh-game playout kernels against the CPU oracle, bit for bit, on many seeds — far more games than the test
suite plays (rare paths: weddings that stay unsolved, counter calls, Black announcements, 12-card hands of one suit ...).

  python tests/soak_parity.py --seconds 60 [--samplers]
Every round draws a new (seed, first_id, epoch), plays 2^20 games on the GPU (rs-full-doko with announcements, rs-full-doko with the
no-announcement policy, rs-doko) and on the oracle (all host threads), and compares points and step counts of every game.
With --samplers every round also compares, on fresh info-states at every stage of a game (oracle_lib.Bulk mode 1) and on BASELINE's
mid-game states (mode 0): 2^19 card_matching samples (hands, hidden reservations, status; every sample is_consistent), 2^18
sample_assignment samples, and the exact point sums of 4096 leaves x 32 rollouts with and without determinization."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch

    import master_doko_reinforcement_learning_b200 as pkg
    import oracle_lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--games", type=int, default=1 << 20)
    ap.add_argument("--samplers", action="store_true")
    a = ap.parse_args()
    dk = pkg.DokoCuda(0)
    L = oracle_lib.load()
    prng = np.random.default_rng(20261018)
    t0 = time.time()
    rounds, games, steps_total, mismatches = 0, 0, 0, 0
    samples_fdo = samples_doko = rollouts = 0
    hist = {}
    while time.time() - t0 < a.seconds:
        seed = int(prng.integers(0, 1 << 62))
        first_id = int(prng.integers(0, 1 << 40))
        epoch = int(prng.integers(0, 1 << 20))
        for engine, ann in ((pkg.DK_FDO, True), (pkg.DK_FDO, False), (pkg.DK_DOKO, False)):
            pts, st = dk.playout(engine, a.games, dk.rng(seed, first_id, epoch), flags=pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS if ann else 0)
            torch.cuda.synchronize()
            ref = oracle_lib.playout_philox(L, engine, a.games, seed, first_id=first_id, epoch=epoch, with_announcements=ann)
            gp, gs = pts.cpu().numpy(), st.cpu().numpy().astype(np.uint32)
            bad = int((gp != ref["points"]).any(axis=1).sum() + (gs != ref["steps"]).sum())
            mismatches += bad
            games += a.games
            steps_total += int(ref["steps"].sum())
            if engine == pkg.DK_FDO and ann:
                for k in (int(ref["steps"].min()), int(ref["steps"].max())):
                    hist[k] = hist.get(k, 0) + 1
        if a.samplers:
            fid = first_id & ((1 << 36) - 1)
            for mode in (0, 1):
                b = oracle_lib.Bulk(L, 1, 8192, seed, first_id=fid, epoch=epoch & 0xFFFF, mode=mode)
                states = torch.from_numpy(b.bytes()).cuda()
                hands, res, status = dk.determinize(1, states, 32, dk.rng(seed, fid, 7, first_sub=3))
                torch.cuda.synchronize()
                h_o, r_o, s_o, cons, _ = b.determinize(32, epoch=7, first_sub=3)
                mismatches += int((status.cpu().numpy() != s_o).sum()) + int((hands.cpu().numpy().astype(np.uint64) != h_o).any(axis=-1).sum())
                mismatches += int((res.cpu().numpy() != r_o).any(axis=-1).sum()) + int((cons != 0).sum())
                samples_fdo += 8192 * 32
                sub = states[:2048].contiguous()
                for det in (True, False):
                    sums = dk.leaf_rollouts(sub, 32, dk.rng(seed, fid, 9), determinize=det)
                    torch.cuda.synchronize()
                    small = oracle_lib.Bulk(L, 1, 2048, seed, first_id=fid, epoch=epoch & 0xFFFF, mode=mode)
                    exp, _ = small.leaf_rollouts(32, epoch=9, determinize=det)
                    mismatches += int((sums.cpu().numpy() != exp).any(axis=-1).sum())
                    rollouts += 2048 * 32
                d = oracle_lib.Bulk(L, 0, 4096, seed, first_id=fid, epoch=epoch & 0xFFFF, mode=mode)
                hands, _, status = dk.determinize(0, torch.from_numpy(d.bytes()).cuda(), 32, dk.rng(seed, fid, 6))
                torch.cuda.synchronize()
                h_o, _, s_o, _, _ = d.determinize(32, epoch=6)
                mismatches += int((status.cpu().numpy() != s_o).sum()) + int((hands.cpu().numpy().astype(np.uint64) != h_o).any(axis=-1).sum())
                samples_doko += 4096 * 32
        rounds += 1
    print(json.dumps({"rounds": rounds, "games_compared": games, "game_steps_compared": steps_total, "mismatching_games": mismatches,
                      "card_matching_samples_compared": samples_fdo, "sample_assignment_samples_compared": samples_doko,
                      "leaf_rollouts_compared": rollouts, "seconds": time.time() - t0, "extreme_step_counts_seen": sorted(hist)}))
    sys.exit(1 if mismatches else 0)


if __name__ == "__main__":
    main()
