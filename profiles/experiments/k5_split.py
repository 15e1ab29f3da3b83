"""Fused env-step+encode kernel vs step-only launch followed by the encode kernel (2^22 games)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 22
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(30):
    dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
snap = states.clone()
act = torch.empty((n,), dtype=torch.uint8, device="cuda")
def timed(fn, iters=6):
    tot = 0.0
    for it in range(iters + 2):
        states.copy_(snap); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: tot += e0.elapsed_time(e1)
    return tot / iters
out = {}
out["fused_ms"] = timed(lambda: dk.step_random_encode(states, dk.rng(SEED, 0, 100), obs_out=obs, action_out=act))
out["step_only_ms"] = timed(lambda: dk.step_random_encode(states, dk.rng(SEED, 0, 100), want_obs=False, action_out=act))
out["encode_only_ms"] = timed(lambda: dk.encode(pkg.DK_LAYOUT_FDO_PI311, states, out=obs))
def two():
    dk.step_random_encode(states, dk.rng(SEED, 0, 100), want_obs=False, action_out=act)
    dk.encode(pkg.DK_LAYOUT_FDO_PI311, states, out=obs)
out["step_then_encode_ms"] = timed(two)
print(json.dumps(out))
