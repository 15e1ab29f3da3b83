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


def pimc_decide(dk, states, n_det, strategy, rng_seed, first_id=0, epoch=0, n_rollouts=None, uct_iterations=None, uct_c=1.4, group=None):
    """Sharded PIMC move decision (DefaultImpiPolicy::execute, compare_impi.rs:212-372): every rank takes a contiguous share of the
    determinizations of EVERY root (sample numbers first_sub .. first_sub+count-1), evaluates them with the flat Monte-Carlo policy
    (`n_rollouts`) or the UCT search (`uct_iterations`), reduces them to integer root statistics, all-reduces those and picks.
    A root without a single successful sample gets the reference's fallback: a random action among the allowed non-announcement
    actions (compare_impi.rs:357-368; dk_random_action on the stream (rng_seed, first_id + i, epoch)) — the same on every rank, since the
    statistics are.  All ranks return (actions uint8 [n] — 0xFF only for finished games —, stats int64 [n, ROOT_STATS])."""
    import torch.distributed as dist

    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_available() and dist.is_initialized() else (1, 0)
    first_sub, count = shard_range(n_det, rank, world)
    allowed = dk.legal_mask(1, states)
    stats = None
    if count > 0:
        rng = dk.rng(rng_seed, first_id, epoch, first_sub)
        if uct_iterations is not None:
            visits, _, _, status = dk.uct_search(states, uct_iterations, uct_c, rng, trees_per_root=count, determinize=True)
        else:
            visits, _, status = dk.pimc_evaluate(states, count, n_rollouts, rng, want_values=False)
        stats = dk.pimc_root_stats(visits, allowed, status)
    else:
        import torch

        from .api import ROOT_STATS
        stats = torch.zeros((states.shape[0], ROOT_STATS), dtype=torch.int64, device=states.device)
    if world > 1:
        dk.allreduce_root_stats(stats)
    import torch

    action = dk.pimc_pick(strategy, stats, allowed)
    fallback = dk.random_action(1, states, dk.rng(rng_seed, first_id, epoch), flags=0)
    return torch.where(action == 0xFF, fallback, action), stats
