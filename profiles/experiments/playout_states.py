"""dk_playout from stored records (one rollout per state) at 2^22 mid-game states, with and without announcements."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import master_doko_reinforcement_learning_b200 as pkg
dk = pkg.DokoCuda(0)
SEED = 0xD0C05EED
n = 1 << 22
out = {}
for adv in (0, 30):
    states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
    for k in range(adv): dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
    pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
    for flags, name in ((pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS, "ann"), (0, "noann")):
        for _ in range(2): dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 0, 2), states=states, flags=flags, points_out=pts, steps_out=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 0, 2), states=states, flags=flags, points_out=pts, steps_out=st)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 / 1e3
        steps = float(st.sum(dtype=torch.int64))
        out[f"adv{adv}_{name}"] = {"ms": t * 1e3, "playouts_per_s": n / t, "game_steps_per_s": steps / t, "mean_steps": steps / n}
print(json.dumps(out))
