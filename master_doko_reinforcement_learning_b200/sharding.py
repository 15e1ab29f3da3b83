"""Host-side sharding of independent units (games / info-states / leaves / rollouts) over ranks.

The path has no data-path collective: rank r owns a contiguous range of unit ids and the Philox counters depend on the unit id only,
so every GPU count produces identical per-unit results.  The only exchange is the integer sum of root statistics.
"""


def shard_range(n_units, rank, world):
    """Contiguous partition of range(n_units): returns (first, count) of `rank`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def leaf_rollout_root_stats(dk, states, rollouts_per_leaf, rng_seed, first_id=0, epoch=0, determinize=True, group=None):
    """Config 4: every rank runs its share of the rollouts of EVERY leaf (rollout numbers first_sub .. first_sub+count-1), then the
    exact integer point sums are all-reduced so that all ranks hold identical root statistics.  Single-process when torch.distributed
    is not initialised."""
    import torch.distributed as dist

    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_available() and dist.is_initialized() else (1, 0)
    first_sub, count = shard_range(rollouts_per_leaf, rank, world)
    sums = dk.leaf_rollouts(states, count, dk.rng(rng_seed, first_id, epoch, first_sub), determinize=determinize)
    if world > 1:
        dk.allreduce_root_stats(sums)
    return sums
