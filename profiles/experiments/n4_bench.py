"""N4 kernels at 2^21 rows: dk_encode_ipi (guessed hands = the true hands of the other seats) and dk_pack_replay_records."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 21
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(30): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
raw = states.view(torch.uint8).reshape(n, 128)
hands = raw[:, :32].contiguous().view(torch.int64).reshape(n, 4).clone()
meta = raw[:, 124:128].contiguous().view(torch.int32).reshape(n)
cur = ((meta >> 2) & 3).to(torch.int64)
hands.scatter_(1, cur.unsqueeze(1), torch.zeros((n, 1), dtype=torch.int64, device="cuda"))      # the observer's own hand is not guessed
res = torch.full((n, 4), 0xFF, dtype=torch.uint8, device="cuda")
nxt = ((cur + 1) & 3).to(torch.uint8)
out = {}
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3
obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
err = torch.empty((n,), dtype=torch.uint8, device="cuda")
try:
    t = timed(lambda: dk.encode_ipi(states, hands, res, nxt, out=obs))
    o, e = dk.encode_ipi(states, hands, res, nxt, out=obs)
    b = 128 + 32 + 4 + 1 + 2488 + 1
    out["encode_ipi_2p21"] = {"ms": t * 1e3, "GBps": n * b / t / 1e9, "frac_of_6549": n * b / t / 1e9 / 6549.4, "errors": int((e != 0).sum())}
except Exception as ex:
    out["encode_ipi_2p21"] = {"error": str(ex)[:200]}
value = torch.rand((n, 4), dtype=torch.float32, device="cuda"); policy = torch.rand((n, 39), dtype=torch.float32, device="cuda")
rec = torch.empty((n, 2684), dtype=torch.uint8, device="cuda")
t = timed(lambda: dk.pack_replay_records(obs, value, policy, out=rec))
b = 2488 + 16 + 156 + 2684
out["pack_replay_records_2p21"] = {"ms": t * 1e3, "GBps": n * b / t / 1e9, "frac_of_6549": n * b / t / 1e9 / 6549.4}
print(json.dumps(out))
