"""K5 (lock-step env step + 311-token observation, 2^22 games): per-thread record loads / stores against TMA record tiles
(default; DOKO_CUDA_NO_TMA=1 selects the per-thread form).  Prints the time and a checksum of states + observations after 8 steps (must be equal for both forms)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 22
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(30): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
act = torch.empty((n,), dtype=torch.uint8, device="cuda")
ctr = [100]
def k5():
    ctr[0] += 1
    dk.step_random_encode(states, dk.rng(SEED, 0, ctr[0]), obs_out=obs, action_out=act)
for _ in range(2): k5()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(6): k5()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 6 / 1e3
w = torch.arange(1, 312, device="cuda", dtype=torch.int64)
print(json.dumps({"tma": not os.environ.get("DOKO_CUDA_NO_TMA"), "ms": t * 1e3, "GBps": n * 2744 / t / 1e9, "frac_of_6549": n * 2744 / t / 1e9 / 6549.4,
                  "obs_checksum": int((obs * w).sum()), "state_checksum": int(states.view(torch.int64).sum()), "act_checksum": int(act.to(torch.int64).sum())}))
