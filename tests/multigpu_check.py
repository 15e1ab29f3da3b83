#!/usr/bin/env python3
"""torchrun --nproc-per-node N tests/multigpu_check.py — on-GPU multi-rank check (run with `gpurun --gpus N`):
sharded leaf rollouts + dk_allreduce_root_stats equal the single-GPU result, and the sharded PIMC decisions (flat Monte-Carlo and UCT per
determinization → root statistics → all-reduce → pick) equal the single-GPU fuse, on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import master_doko_reinforcement_learning_b200 as pkg
    from master_doko_reinforcement_learning_b200.sharding import leaf_rollout_root_stats, pimc_decide

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dk = pkg.DokoCuda(local)
    dk.comm_init()
    seed, n_leaves, R = 0xD0C05EED, 1024, 1024
    states = dk.new_games(pkg.DK_FDO, n_leaves, dk.rng(seed, 0, 0))
    for k in range(20):
        dk.step_random_encode(states, dk.rng(seed, 0, k), want_obs=False)
    sums = leaf_rollout_root_stats(dk, states, R, seed, first_id=0, epoch=7)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    buf = torch.zeros((n_leaves, 39), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(20):
        dk.allreduce_root_stats(buf)
    ev1.record()
    torch.cuda.synchronize()
    full = dk.leaf_rollouts(states, R, dk.rng(seed, 0, 7), determinize=True)      # all rollouts on this GPU
    torch.cuda.synchronize()
    ok = torch.equal(sums, full)
    # sharded PIMC decisions (flat MC and UCT per determinization) == the single-GPU fuse over all determinizations
    roots, n_det = states[:256], 16
    allowed = dk.legal_mask(pkg.DK_FDO, roots)
    for strategy in (pkg.FUSE_MAX_N, pkg.FUSE_AVERAGE):
        act, _ = pimc_decide(dk, roots, n_det, strategy, seed, first_id=0, epoch=8, n_rollouts=32)
        v, _, st = dk.pimc_evaluate(roots, n_det, 32, dk.rng(seed, 0, 8), want_values=False)
        f = dk.fuse(strategy, v, allowed, st)[0]
        ok = ok and torch.equal(torch.where(f == 0xFF, act, f), act)      # (a root without a successful sample gets pimc_decide's random fallback)
    act, _ = pimc_decide(dk, roots, n_det, pkg.FUSE_MAX_N, seed, first_id=0, epoch=9, uct_iterations=64)
    v, _, _, st = dk.uct_search(roots, 64, 1.4, dk.rng(seed, 0, 9), trees_per_root=n_det, determinize=True)
    f = dk.fuse(pkg.FUSE_MAX_N, v, allowed, st)[0]
    ok = ok and torch.equal(torch.where(f == 0xFF, act, f), act)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MULTIGPU world={world} root-stats identical on all ranks: {bool(flag.item())}; all-reduce of {n_leaves}x39 i64 = "
              f"{ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us")
    dk.comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
