"""rs-doko observation encoders (110 / 114 tokens) at 2^22 states: time and HBM fraction."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 22
states = dk.new_games(pkg.DK_DOKO, n, dk.rng(SEED, 0, 0))
for k in range(24):
    m = dk.legal_mask(pkg.DK_DOKO, states)
    act = (torch.log2((m & -m).to(torch.float64)) + 0.5).to(torch.uint8)
    assert int(dk.apply(pkg.DK_DOKO, states, act).sum()) == 0
out = {}
for layout, L, name in ((pkg.DK_LAYOUT_DO110, 110, "do110"), (pkg.DK_LAYOUT_DO114, 114, "do114")):
    obs = torch.empty((n, L), dtype=torch.int64, device="cuda")
    for _ in range(2): dk.encode(layout, states, out=obs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dk.encode(layout, states, out=obs)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 / 1e3
    out[name] = {"ms": t * 1e3, "GBps": n * (128 + 8 * L) / t / 1e9, "frac_of_6549": n * (128 + 8 * L) / t / 1e9 / 6549.4, "checksum": int(obs.sum())}
    del obs
print(json.dumps(out))
