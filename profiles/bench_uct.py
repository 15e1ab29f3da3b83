"""UCT search throughput vs number of resident trees and iterations (one B200).  Prints one JSON line."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 0xD0C05EED


def main():
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    dk = pkg.DokoCuda(0)
    n_max = 8192
    states = dk.new_games(pkg.DK_FDO, n_max, dk.rng(SEED, 0, 0))
    # advance to mixed mid-game positions with the lock-step random step (card_index 0..40)
    for step in range(30):
        dk.step_random_encode(states[: n_max // 2], dk.rng(SEED, 0, step), flags=0, want_obs=False)
    out = {}
    for n, T, iters in ((2048, 8, 256), (2048, 64, 256), (8192, 64, 256), (8192, 32, 1024), (8192, 128, 64)):
        need = dk.L.dk_uct_workspace_bytes(n * T, iters)
        ws = torch.empty((need // 8 + 2,), dtype=torch.int64, device="cuda")
        roots = states[:n]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dk.uct_search(roots, iters, 1.4, dk.rng(SEED, 0, 13), trees_per_root=T, determinize=True, workspace=ws)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(2):
            dk.uct_search(roots, iters, 1.4, dk.rng(SEED, 0, 13), trees_per_root=T, determinize=True, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 2 / 1e3
        out[f"{n}x{T}x{iters}"] = {"sec": t, "trees": n * T, "iterations_per_s": n * T * iters / t, "decisions_per_s": n / t, "workspace_GB": need / 1e9}
        del ws
    print(json.dumps(out))


if __name__ == "__main__":
    main()
