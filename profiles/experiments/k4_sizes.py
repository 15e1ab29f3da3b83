"""K4 (leaf rollouts, 1024 per leaf) at small leaf counts: one block per leaf (DOKO_CUDA_NO_SPLIT=1) against several blocks per leaf."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
R = 1024
states = dk.new_games(pkg.DK_FDO, 1024, dk.rng(SEED, 0, 5))
for k in range(30 + 12): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3
out = {}
for det in (False, True):
    for n in (1, 16, 128, 512, 1024):
        sums = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        t = timed(lambda: dk.leaf_rollouts(states[:n], R, dk.rng(SEED, 0, 11), determinize=det, out=sums))
        out[f"det{int(det)}_{n}"] = {"ms": t * 1e3, "rollouts_per_s": n * R / t, "checksum": int(sums.sum()), "abs_checksum": int(sums.abs().sum())}
print(json.dumps(out))
