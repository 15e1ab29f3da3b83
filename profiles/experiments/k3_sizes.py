"""K3 (full-rules determinizer) at several info-state counts, 4096 samples each: wave quantisation of the block-per-info-state grid
(DOKO_CUDA_NO_SPLIT=1) against the grid that gives small batches several blocks per info-state."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
S = 4096
NMAX = 1 << 16
states = dk.new_games(pkg.DK_FDO, NMAX, dk.rng(SEED, 0, 5))
for k in range(30 + 12): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
hands = torch.empty((NMAX, S, 4), dtype=torch.int64, device="cuda")
res = torch.empty((NMAX, S, 4), dtype=torch.uint8, device="cuda")
status = torch.empty((NMAX, S), dtype=torch.uint8, device="cuda")
def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3
out = {}
for n_info in (1, 64, 1024, 4096, 5328, 8192, 16384, 65536):
    def k3():
        dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_FDO, n_info, S, pkg.api._ptr(states), ctypes.byref(dk.rng(SEED, 0, 9)), pkg.api._ptr(hands),
                                      pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")
    t = timed(k3)
    out[str(n_info)] = {"ms": t * 1e3, "samples_per_s": n_info * S / t, "out_GBps": n_info * S * 37 / t / 1e9}
print(json.dumps(out))
