"""Per-call device time of one lock-step self-play turn (2^20 games), CUDA events around each C-ABI call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
SEED = 0xD0C05EED
dk = pkg.DokoCuda(0)
n = 1 << 20
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 0))
sp = dk.self_play(n, n * 4)
for t in range(12):
    sp.reset(); sp.begin_turn(states, 0, 1.0, dk.rng(SEED, 0, t)); sp.uniform_search(dk.rng(SEED, 0, t)); sp.end_turn(states)
snap = states.clone()
acc = {"begin_turn(plan+scan+encode)": 0.0, "uniform_search": 0.0, "end_turn(apply)": 0.0}
iters = 6
for it in range(iters + 2):
    states.copy_(snap); sp.reset(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); sp.begin_turn(states, 0, 1.0, dk.rng(SEED, 0, 12)); ev[1].record(); sp.uniform_search(dk.rng(SEED, 0, 12)); ev[2].record(); sp.end_turn(states); ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        for k, (a, b) in zip(acc, ((0, 1), (1, 2), (2, 3))):
            acc[k] += ev[a].elapsed_time(ev[b]) / iters
print(json.dumps({k: round(v * 1e3, 1) for k, v in acc.items()}), "us")
