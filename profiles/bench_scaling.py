#!/usr/bin/env python3
"""Weak-scaling sweep of BASELINE configs 2-5 over the GPUs of one box: one process per GPU, units sharded by id, no
data-path collective except config 4's integer all-reduce of the root statistics.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/bench_scaling.py

Per-rank work is fixed (config 2: 2^24 games, config 3: 4096 info-states x 4096 samples, config 4: 1024 leaves x 1024 rollouts + the
all-reduce over all 1024 N leaves, config 5: 2^22 games); every number is CUDA-event time, max over ranks; rank 0 prints one JSON line.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 0xD0C05EED


def main():
    import torch
    import torch.distributed as dist

    import master_doko_reinforcement_learning_b200 as pkg

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dk = pkg.DokoCuda(local)
    if world > 1:
        dk.comm_init()

    def timed(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters / 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    out = {"n_gpus": world, "scaling": "weak"}
    # config 2: 2^24 fresh full-rules playouts per GPU, game ids [rank * 2^24, (rank + 1) * 2^24)
    n = 1 << 24
    pts = torch.empty((n, 4), dtype=torch.int32, device="cuda"); st = torch.empty((n,), dtype=torch.int32, device="cuda")
    t = timed(lambda: dk.playout(pkg.DK_FDO, n, dk.rng(SEED, rank * n, 2), flags=pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS, points_out=pts, steps_out=st), iters=3)
    steps = total(int(st.sum(dtype=torch.int64)))
    out["config2_fdo_playouts"] = {"sec": t, "games_per_gpu": n, "game_steps_per_s": steps / t, "games_per_s": n * world / t}
    del pts, st
    # config 5: lock-step env step + 311-token encode, 2^22 games per GPU
    n5 = 1 << 22
    states = dk.new_games(pkg.DK_FDO, n5, dk.rng(SEED, rank * n5, 5))
    for k in range(30):
        dk.step_random_encode(states, dk.rng(SEED, rank * n5, k), want_obs=False)
    obs = torch.empty((n5, 311), dtype=torch.int64, device="cuda")
    act = torch.empty((n5,), dtype=torch.uint8, device="cuda")
    ctr = [100]

    def k5():
        ctr[0] += 1
        dk.step_random_encode(states, dk.rng(SEED, rank * n5, ctr[0]), obs_out=obs, action_out=act)

    t = timed(k5)
    out["config5_step_encode"] = {"sec": t, "games_per_gpu": n5, "step_encodes_per_s": n5 * world / t, "algorithmic_GBps_per_gpu": n5 * 2744 / t / 1e9}
    del obs
    # config 3: 4096 info-states x 4096 samples per GPU
    n_info, S = 4096, 4096
    sub = states[:n_info].clone()
    for k in range(12):
        dk.step_random_encode(sub, dk.rng(SEED, rank * n5, 500 + k), want_obs=False)
    hands = torch.empty((n_info, S, 4), dtype=torch.int64, device="cuda")
    res = torch.empty((n_info, S, 4), dtype=torch.uint8, device="cuda")
    status = torch.empty((n_info, S), dtype=torch.uint8, device="cuda")
    import ctypes

    def k3():
        dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_FDO, n_info, S, pkg.api._ptr(sub), ctypes.byref(dk.rng(SEED, rank * n_info, 9)), pkg.api._ptr(hands),
                                      pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

    t = timed(k3, iters=3)
    out["config3_determinize"] = {"sec": t, "samples_per_s": n_info * S * world / t, "dead_ends": total(int((status != 0).sum()))}
    del hands, res, status
    # config 4: 1024 leaves x 1024 rollouts per GPU, then ONE integer all-reduce so that every rank holds the reward sums of all leaves
    n_leaves, R = 1024, 1024
    leaves = sub[:n_leaves]
    all_sums = torch.zeros((n_leaves * world, 4), dtype=torch.int64, device="cuda")
    mine = all_sums[rank * n_leaves:(rank + 1) * n_leaves]

    def k4():
        all_sums.zero_()
        dk.leaf_rollouts(leaves, R, dk.rng(SEED, rank * n_leaves, 11), determinize=True, out=mine)
        if world > 1:
            dk.allreduce_root_stats(all_sums)

    t = timed(k4)
    chk = torch.tensor([int(all_sums.sum())], device="cuda", dtype=torch.int64)
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).item())
    out["config4_leaf_rollouts_allreduce"] = {"sec": t, "rollouts_per_s": n_leaves * R * world / t, "leaves_total": n_leaves * world,
                                              "root_stats_identical_on_all_ranks": same}
    if world > 1:
        ref_sums = all_sums.clone()
        t_ar = timed(lambda: dk.allreduce_root_stats(all_sums), iters=20)
        out["config4_allreduce_only_us"] = t_ar * 1e6
        # the same work as a two-stream pipeline: the all-reduce of batch k (stream B) runs under the rollouts of batch k + 1 (stream A);
        # two result buffers, events order "rollouts done -> reduce" and "reduce done -> buffer reusable"
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        bufs = [torch.zeros((n_leaves * world, 4), dtype=torch.int64, device="cuda") for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in free:
            e.record(sb)
        batch = [0]

        def k4_pipelined():
            b = batch[0] & 1
            batch[0] += 1
            sa.wait_event(free[b])
            with torch.cuda.stream(sa):
                bufs[b].zero_()
                dk.leaf_rollouts(leaves, R, dk.rng(SEED, rank * n_leaves, 11), determinize=True, out=bufs[b][rank * n_leaves:(rank + 1) * n_leaves],
                                 stream=sa.cuda_stream)
                done[b].record(sa)
            sb.wait_event(done[b])
            dk.allreduce_root_stats(bufs[b], stream=sb.cuda_stream)
            free[b].record(sb)

        def run_pipelined(iters):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                k4_pipelined()
            torch.cuda.current_stream().wait_stream(sa)
            torch.cuda.current_stream().wait_stream(sb)
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / iters / 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        run_pipelined(4)
        tp = run_pipelined(20)
        same_p = bool(torch.equal(bufs[0], ref_sums) and torch.equal(bufs[1], ref_sums))
        out["config4_pipelined"] = {"sec_per_batch": tp, "rollouts_per_s": n_leaves * R * world / tp, "equals_unpipelined_result": same_p}
    out["launches_rank0"] = dk.launch_count()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dk.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
