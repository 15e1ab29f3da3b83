#!/usr/bin/env python3
"""Per-kernel timings (CUDA events, warm, inputs > L2 where it matters) for K1..K5; prints one JSON object."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 0xD0C05EED


def timed(fn, iters=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3


def main():
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    dk = pkg.DokoCuda(0)
    out = {}
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {}
    hbm = peaks.get("hbm_gbs", 6650.0)

    # K1 rs-doko fresh playouts, 1M games (config 1)
    n = 1_000_000
    pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
    t = timed(lambda: dk.playout(pkg.DK_DOKO, n, dk.rng(SEED, 0, 1), points_out=pts, steps_out=st))
    out["K1_doko_playout_1M"] = {"sec": t, "games_per_s": n / t, "game_steps_per_s": 52 * n / t}

    # K2 fdo fresh playouts 2^24 (config 2), both policies
    n = int((1 << 24) * a.scale)
    pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
    for name, fl in (("with_announcements", 1), ("no_announcement_policy", 0)):
        t = timed(lambda: dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 0, 2), flags=fl, points_out=pts, steps_out=st), iters=3)
        steps = int(st.sum(dtype=torch.int64))
        out[f"K2_fdo_playout_2p24_{name}"] = {"sec": t, "games_per_s": n / t, "game_steps_per_s": steps / t, "mean_steps": steps / n}
    del pts, st

    # K5 lock-step step + encode, 2^22 games (config 5): states mid-game
    n = int((1 << 22) * a.scale)
    states = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
    for k in range(30):
        dk.step_random_encode(states, dk.rng(SEED, 0, k), want_obs=False)
    obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
    act = torch.empty((n,), dtype=torch.uint8, device="cuda")
    ctr = [100]

    def k5():
        ctr[0] += 1
        dk.step_random_encode(states, dk.rng(SEED, 0, ctr[0]), obs_out=obs, action_out=act)

    t = timed(k5, iters=5)
    out["K5_step_encode_2p22"] = {"sec": t, "step_encodes_per_s": n / t, "algorithmic_GBps": n * 2744 / t / 1e9, "hbm_frac_of_measured": n * 2744 / t / 1e9 / hbm,
                                  "bytes_per_unit": 2744}
    t = timed(lambda: dk.encode(pkg.DK_LAYOUT_FDO_PI311, states, out=obs), iters=5)
    out["encode_pi311_2p22"] = {"sec": t, "encodes_per_s": n / t, "algorithmic_GBps": n * 2616 / t / 1e9, "hbm_frac_of_measured": n * 2616 / t / 1e9 / hbm}
    ms = torch.empty((n,), dtype=torch.int64, device="cuda")
    t = timed(lambda: dk.legal_mask(pkg.DK_FDO, states, out=ms), iters=5)
    out["legal_mask_2p22"] = {"sec": t, "states_per_s": n / t, "algorithmic_GBps": n * 136 / t / 1e9}
    del obs

    # K3 determinization: 64K info-states x 4096 samples is 2.7e8 samples = 10 GB of hands; run 4096 states x 4096 samples and scale
    n_info = int(4096 * a.scale); S = 4096
    sub = states[:n_info].clone()
    for k in range(12):
        dk.step_random_encode(sub, dk.rng(SEED, 0, 500 + k), want_obs=False)
    hands = torch.empty((n_info, S, 4), dtype=torch.int64, device="cuda")
    res = torch.empty((n_info, S, 4), dtype=torch.uint8, device="cuda")
    status = torch.empty((n_info, S), dtype=torch.uint8, device="cuda")

    def k3():
        dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_FDO, n_info, S, pkg.api._ptr(sub), __import__("ctypes").byref(dk.rng(SEED, 0, 9)), pkg.api._ptr(hands),
                                      pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

    t = timed(k3, iters=3)
    out["K3_determinize_4096x4096"] = {"sec": t, "samples_per_s": n_info * S / t, "dead_ends": int((status != 0).sum()), "out_GBps": n_info * S * 37 / t / 1e9}
    # K3 (rs-doko): sample_assignment on simplified-engine info-states, 4096 states x 1024 samples
    d_states = dk.new_games(pkg.DK_DOKO, n_info, dk.rng(SEED, 0, 0))
    for k in range(20):                                       # 4 reservations + 16 cards: lowest legal action each turn
        m = dk.legal_mask(pkg.DK_DOKO, d_states)
        act = ((torch.log2((m & -m).to(torch.float64)) + 0.5).to(torch.uint8))
        assert int(dk.apply(pkg.DK_DOKO, d_states, act).sum()) == 0
    S2 = 1024

    def k3d():
        dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_DOKO, n_info, S2, pkg.api._ptr(d_states), __import__("ctypes").byref(dk.rng(SEED, 0, 9)), pkg.api._ptr(hands),
                                      pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

    t = timed(k3d, iters=3)
    out["K3_sample_assignment_4096x1024"] = {"sec": t, "samples_per_s": n_info * S2 / t, "dead_ends": int((status.view(-1)[: n_info * S2] != 0).sum())}
    del hands, res, status

    # K4 leaf rollouts: 1024 leaves x 1024 rollouts (per-GPU share of config 4)
    n_leaves, R = int(1024 * a.scale), 1024
    leaves = sub[:n_leaves]
    sums = torch.empty((n_leaves, 4), dtype=torch.int64, device="cuda")
    for det in (True, False):
        t = timed(lambda: dk.leaf_rollouts(leaves, R, dk.rng(SEED, 0, 11), determinize=det, out=sums), iters=3)
        out[f"K4_leaf_rollouts_1024x1024_det{int(det)}"] = {"sec": t, "rollouts_per_s": n_leaves * R / t}
    # N2 PIMC decision: 1024 roots x 64 determinizations x 32 rollouts x (legal actions), evaluate + fuse
    n_roots, n_det, R2 = int(1024 * a.scale), 64, 32
    roots = sub[:n_roots]
    allowed = dk.legal_mask(pkg.DK_FDO, roots)
    n_legal = float(sum(bin(int(m) & ((1 << 39) - 1)).count("1") for m in allowed.cpu().tolist())) / n_roots

    def n2():
        v, _, st = dk.pimc_evaluate(roots, n_det, R2, dk.rng(SEED, 0, 12), want_values=False)
        dk.fuse(pkg.FUSE_MAX_N, v, allowed, st)

    t = timed(n2, iters=3)
    out["N2_pimc_decision_1024x64x32"] = {"sec": t, "decisions_per_s": n_roots / t, "mean_legal_actions": n_legal,
                                          "rollouts_per_s": n_roots * n_det * R2 * n_legal / t}
    # N3 UCT search: 2048 roots x 8 determinized trees x 256 iterations (mcts_cap_* per-sample search)
    n_uct, T, iters_uct = int(2048 * a.scale), 8, 256
    uct_roots = sub[:n_uct]
    need = dk.L.dk_uct_workspace_bytes(n_uct * T, iters_uct)
    ws = torch.empty((need // 8 + 2,), dtype=torch.int64, device="cuda")
    t = timed(lambda: dk.uct_search(uct_roots, iters_uct, 1.4, dk.rng(SEED, 0, 13), trees_per_root=T, determinize=True, workspace=ws), iters=3)
    out["N3_uct_search_2048x8x256"] = {"sec": t, "trees": n_uct * T, "iterations_per_s": n_uct * T * iters_uct / t, "decisions_per_s": n_uct / t,
                                       "workspace_GB": need / 1e9}
    del ws

    # N1 self-play driver: 2^20 games in lock-step, one full turn = plan + scan + encode-into-row + stand-in search + apply
    n_sp = int((1 << 20) * a.scale)
    sp_states = dk.new_games(pkg.DK_FDO, n_sp, dk.rng(SEED, 0, 0))
    sp = dk.self_play(n_sp, n_sp * 4)
    for t in range(12):                                   # into the card phase, where most turns are searched (not forced)
        sp.reset()
        sp.begin_turn(sp_states, 0, 1.0, dk.rng(SEED, 0, t))
        sp.uniform_search(dk.rng(SEED, 0, t))
        sp.end_turn(sp_states)
    snapshot = sp_states.clone()

    def n1():
        sp_states.copy_(snapshot)
        sp.reset()
        sp.begin_turn(sp_states, 0, 1.0, dk.rng(SEED, 0, 12))
        sp.uniform_search(dk.rng(SEED, 0, 12))
        sp.end_turn(sp_states)

    def n1_copy_only():
        sp_states.copy_(snapshot)

    t_all, t_copy = timed(n1, iters=5), timed(n1_copy_only, iters=5)
    t = t_all - t_copy
    rows, dropped, _ = sp.counts()
    # bytes per game-turn: plan 128 r + 9 w; encode 128 r + 2488 w + 5 w (+156 w forced); search 9 r + 157 w; apply 128 r + 128 w + 9 r + 156 r + 156 w
    bytes_turn = 128 + 9 + 128 + 2488 + 5 + 9 + 157 + 128 + 128 + 9 + 156 + 156
    out["N1_selfplay_turn_2p20"] = {"sec": t, "game_turns_per_s": n_sp / t, "rows": rows, "dropped": dropped, "bytes_per_game_turn": bytes_turn,
                                    "algorithmic_GBps": n_sp * bytes_turn / t / 1e9, "hbm_frac_of_measured": n_sp * bytes_turn / t / 1e9 / hbm}
    out["launches"] = dk.launch_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
