"""Times the fresh full-rules playout kernel (K2) alone: python tools/k2_time.py [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
n = 1 << 24
pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
out = {}
for name, fl in (("ann", 1), ("noann", 0)):
    ts = []
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
        for _ in range(2): dk.playout(pkg.DK_FDO, n, dk.rng(7, 0, 2), flags=fl, points_out=pts, steps_out=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dk.playout(pkg.DK_FDO, n, dk.rng(7, 0, 2), flags=fl, points_out=pts, steps_out=st)
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1) / 5, 4))
    out[name] = ts
print(json.dumps(out))
