"""K2 (2^24 fresh full-rules playouts) launch time for the grid cap given in DK_PLAYOUT_GRID_PER_SM (0 = one block per batch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
n = 1 << 24
out = {"grid_per_sm": os.environ.get("DK_PLAYOUT_GRID_PER_SM")}
for flags, name in ((pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS, "with_ann_ms"), (0, "no_ann_ms")):
    pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); steps = torch.empty((n,), dtype=torch.int32, device="cuda")
    for it in range(2): dk.playout(pkg.DK_FDO, n, dk.rng(0xD0C05EED, 0, 2), flags=flags, points_out=pts, steps_out=steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5): dk.playout(pkg.DK_FDO, n, dk.rng(0xD0C05EED, 0, 2), flags=flags, points_out=pts, steps_out=steps)
    e1.record(); torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / 5
    out[name.replace("_ms", "_steps_per_s")] = float(steps.sum(dtype=torch.int64)) / (out[name] * 1e-3)
    out[name.replace("_ms", "_checksum")] = int(pts.sum(dtype=torch.int64)) * 1000003 + int(steps.sum(dtype=torch.int64))
print(json.dumps(out))
