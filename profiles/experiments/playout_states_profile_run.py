"""dk_playout from 2^20 stored mid-game records (no-announcement policy = McEnvState::random_rollout) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 20
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(30): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
for _ in range(3): dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 0, 2), states=states, flags=0, points_out=pts, steps_out=st)
torch.cuda.synchronize()
print("steps", int(st.sum(dtype=torch.int64)))
