"""K4 (1024 leaves x 1024 rollouts, no determinization) launch time for the library in DOKO_CUDA_LIB."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1024
states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
for k in range(42): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
sums = torch.empty((n, 4), dtype=torch.int64, device="cuda")
out = {"lib": os.path.basename(os.environ.get("DOKO_CUDA_LIB", "default"))}
for det in (False, True):
    for it in range(3): dk.leaf_rollouts(states, 1024, dk.rng(SEED, 0, 11), determinize=det, out=sums)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10): dk.leaf_rollouts(states, 1024, dk.rng(SEED, 0, 11), determinize=det, out=sums)
    e1.record(); torch.cuda.synchronize()
    out[f"det{int(det)}_ms"] = e0.elapsed_time(e1) / 10
    out[f"det{int(det)}_checksum"] = int(sums.sum())
print(json.dumps(out))
