"""rs-doko-assignment sampler launches for ncu / timing: 4096 info-states x 1024 samples."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n_info, S = 4096, 1024
d_states = dk.new_games(pkg.DK_DOKO, n_info, dk.rng(SEED, 0, 0))
for k in range(20):
    m = dk.legal_mask(pkg.DK_DOKO, d_states)
    act = (torch.log2((m & -m).to(torch.float64)) + 0.5).to(torch.uint8)
    assert int(dk.apply(pkg.DK_DOKO, d_states, act).sum()) == 0
hands = torch.empty((n_info, S, 4), dtype=torch.int64, device="cuda")
res = torch.empty((n_info, S, 4), dtype=torch.uint8, device="cuda")
status = torch.empty((n_info, S), dtype=torch.uint8, device="cuda")
def run():
    dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_DOKO, n_info, S, pkg.api._ptr(d_states), ctypes.byref(dk.rng(SEED, 0, 9)), pkg.api._ptr(hands),
                                  pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5 / 1e3
print(json.dumps({"ms": t * 1e3, "samples_per_s": n_info * S / t, "dead_ends": int((status != 0).sum()), "checksum": int(hands.sum())}))
