"""One UCT launch for ncu: 2048 roots x 8 determinized trees x 128 iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 2048
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 0))
for step in range(30): dk.step_random_encode(states[: n // 2], dk.rng(SEED, 0, step), flags=0, want_obs=False)
for _ in range(2):
    dk.uct_search(states, 128, 1.4, dk.rng(SEED, 0, 13), trees_per_root=8, determinize=True)
torch.cuda.synchronize()
print("ok")
