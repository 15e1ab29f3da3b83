#!/usr/bin/env python3
"""Touches every kernel with small, ragged sizes — the target of the compute-sanitizer memcheck run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 7


def main():
    import numpy as np
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    dk = pkg.DokoCuda(0)
    for n in (1, 33, 1000, 4099):
        for eng in (pkg.DK_DOKO, pkg.DK_FDO):
            dk.playout(eng, n, dk.rng(SEED, 0, 0), flags=1)
            dk.playout(eng, n, dk.rng(SEED, 0, 0), flags=0)
            st = dk.new_games(eng, n, dk.rng(SEED, 5, 1))
            for k in range(60):
                m = dk.legal_mask(eng, st)
                # lowest legal action
                a = torch.where(m != 0, (m & -m).to(torch.float64).log2().round().to(torch.uint8), torch.zeros_like(m, dtype=torch.uint8))
                dk.apply(eng, st, a, flags=k & 1)
                if k % 13 == 0:
                    dk.playout(eng, n, dk.rng(SEED, 9, k), states=st, flags=k & 1)
                    dk.determinize(eng, st, 5, dk.rng(SEED, 3, k))
                    if eng == pkg.DK_FDO:
                        dk.encode(pkg.DK_LAYOUT_FDO_PI311, st)
                        dk.encode(pkg.DK_LAYOUT_FDO_PI311, st, row_stride=312)
                        dk.leaf_rollouts(st[: min(n, 64)], 37, dk.rng(SEED, 1, k), determinize=True)
                        dk.leaf_rollouts(st[: min(n, 64)], 37, dk.rng(SEED, 1, k), determinize=False)
                    else:
                        dk.encode(pkg.DK_LAYOUT_DO110, st)
                        dk.encode(pkg.DK_LAYOUT_DO114, st)
            dk.terminal(eng, st)
        st = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 2))
        for k in range(110):
            dk.step_random_encode(st, dk.rng(SEED, 0, 10 + k), flags=(1, 0, 0x100, 0x101)[k & 3], want_obs=(k % 7 == 0))
        dk.playout_trace(pkg.DK_DOKO, n, dk.rng(SEED, 0, 3))
        dk.playout_host(pkg.DK_FDO, n, dk.rng(SEED, 0, 4), flags=1)
        dk.playout_host_compact(pkg.DK_FDO, n, dk.rng(SEED, 0, 4), flags=1)
        hands = torch.randint(0, 1 << 40, (n, 4), dtype=torch.int64, device="cuda")
        dk.from_deals(pkg.DK_FDO, hands, torch.zeros(n, dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    print("sanitize smoke done, launches:", dk.launch_count())


if __name__ == "__main__":
    main()
