"""HBM fraction of the record-level operations at 2^22 states (dk_new_games, dk_legal_mask, dk_apply, dk_terminal)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 22
PEAK = 6549.4
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3
out = {}
states = dk.alloc_states(n)
t = timed(lambda: dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5), out=states))
out["new_games"] = {"ms": t * 1e3, "GBps": n * 128 / t / 1e9, "frac": n * 128 / t / 1e9 / PEAK, "games_per_s": n / t}
d_hands = states.view(torch.uint8).reshape(n, 128)[:, :32].contiguous().view(torch.int64).reshape(n, 4)
d_start = ((states.view(torch.uint8).reshape(n, 128)[:, 124] >> 4) & 3).contiguous()
st2 = dk.alloc_states(n)
t = timed(lambda: dk.from_deals(pkg.DK_FDO, d_hands, d_start, out=st2))
out["from_deals"] = {"ms": t * 1e3, "GBps": n * 161 / t / 1e9, "frac": n * 161 / t / 1e9 / PEAK, "equal_to_new_games": bool(torch.equal(st2, states))}
del st2
for k in range(30): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
mask = torch.empty((n,), dtype=torch.int64, device="cuda")
t = timed(lambda: dk.legal_mask(pkg.DK_FDO, states, out=mask))
out["legal_mask"] = {"ms": t * 1e3, "GBps": n * 136 / t / 1e9, "frac": n * 136 / t / 1e9 / PEAK}
act = (torch.log2((mask & -mask).to(torch.float64)) + 0.5).to(torch.uint8)      # the lowest legal action of every state
snap = states.clone()
err = torch.empty((n,), dtype=torch.uint8, device="cuda")
def ap():
    dk.apply(pkg.DK_FDO, states, act, err_out=err)
t_copy = timed(lambda: states.copy_(snap))
def ap2():
    states.copy_(snap); ap()
t = timed(ap2) - t_copy
out["apply"] = {"ms": t * 1e3, "GBps": n * 258 / t / 1e9, "frac": n * 258 / t / 1e9 / PEAK, "errors": int(err.sum())}
t = timed(lambda: dk.terminal(pkg.DK_FDO, states))
out["terminal"] = {"ms": t * 1e3, "GBps_line_granularity": n * (32 + 17) / t / 1e9}
print(json.dumps(out))
