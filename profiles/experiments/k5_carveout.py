import sys, os, json
sys.path.insert(0, "/root/repo")
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
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
t_enc = timed(lambda: dk.encode(pkg.DK_LAYOUT_FDO_PI311, states, out=obs))
k = [100]
def step():
    k[0] += 1
    dk.step_random_encode(states, dk.rng(SEED, 0, k[0] % 8 + 100), obs_out=obs)
def cp(): states.copy_(snap)
t_step = timed(step)
print(json.dumps({"carveout": os.environ.get("DK_ENC_CARVEOUT"), "encode_ms": t_enc, "step_encode_ms": t_step}))
