"""PIMC decision time (1024 roots x 64 determinizations x 32 rollouts) for the library in DOKO_CUDA_LIB."""
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
out = {"lib": os.path.basename(os.environ.get("DOKO_CUDA_LIB", "default"))}
for n_det, R in ((64, 32), (16, 128), (128, 8)):
    def run():
        v, _, st = dk.pimc_evaluate(states, n_det, R, dk.rng(SEED, 0, 12), want_values=False)
        return dk.fuse(pkg.FUSE_MAX_N, v, allowed, st)[0], v
    for _ in range(2): act, v = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): act, v = run()
    e1.record(); torch.cuda.synchronize()
    out[f"{n_det}x{R}"] = {"ms": e0.elapsed_time(e1) / 3, "checksum": int((v.to(torch.int64) * torch.arange(39, device="cuda")).sum()) * 31 + int(act.to(torch.int64).sum())}
print(json.dumps(out))
