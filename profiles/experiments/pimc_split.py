"""dk_pimc_evaluate and dk_fuse timed separately (1024 roots)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1024
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(42): dk.step_random_encode(states, dk.rng(SEED, 0, k if k < 30 else 470 + k), want_obs=False)
allowed = dk.legal_mask(pkg.DK_FDO, states)
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {}
for n_det, R in ((64, 32), (16, 128), (128, 8)):
    v, _, st = dk.pimc_evaluate(states, n_det, R, dk.rng(SEED, 0, 12), want_values=False)
    t_eval = timed(lambda: dk.pimc_evaluate(states, n_det, R, dk.rng(SEED, 0, 12), want_values=False))
    t_fuse = timed(lambda: dk.fuse(pkg.FUSE_MAX_N, v, allowed, st))
    t_fuse_avg = timed(lambda: dk.fuse(pkg.FUSE_AVERAGE, v, allowed, st))
    out[f"{n_det}x{R}"] = {"evaluate_ms": t_eval, "fuse_maxn_ms": t_fuse, "fuse_avg_ms": t_fuse_avg}
print(json.dumps(out))
