"""UCT throughput for the library in DOKO_CUDA_LIB at three tree counts."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n_max = 8192
states = dk.new_games(pkg.DK_FDO, n_max, dk.rng(SEED, 0, 0))
for step in range(30): dk.step_random_encode(states[: n_max // 2], dk.rng(SEED, 0, step), flags=0, want_obs=False)
out = {"lib": os.path.basename(os.environ.get("DOKO_CUDA_LIB", "default"))}
for n, T, iters in ((2048, 8, 256), (2048, 64, 256), (8192, 128, 64)):
    need = dk.L.dk_uct_workspace_bytes(n * T, iters)
    ws = torch.empty((need // 8 + 2,), dtype=torch.int64, device="cuda")
    roots = states[:n]
    v = dk.uct_search(roots, iters, 1.4, dk.rng(SEED, 0, 13), trees_per_root=T, determinize=True, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): v = dk.uct_search(roots, iters, 1.4, dk.rng(SEED, 0, 13), trees_per_root=T, determinize=True, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 2 / 1e3
    out[f"{n}x{T}x{iters}"] = {"ms": t * 1e3, "Miter_per_s": n * T * iters / t / 1e6, "checksum": int((v[0].to(torch.int64) * torch.arange(39, device="cuda")).sum())}
    del ws
print(json.dumps(out))
