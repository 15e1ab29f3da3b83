"""Host-side multi-rank logic on CPU (gloo, world_size 2): contiguous sharding covers every unit exactly once, and summing the
per-rank integer root statistics (computed here by the oracle as a stand-in for the kernel) equals the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

from master_doko_reinforcement_learning_b200.sharding import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib
    from oracle_lib import Fdo

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = oracle_lib.load()
    seed, n_leaves, R = 99, 6, 40
    leaves = []
    for i in range(n_leaves):
        o = Fdo.new_game_philox(L, seed, i, 0)
        for k in range(10 + 5 * i):
            o.random_step(seed, i, 0, True, k)
        leaves.append(o)
    first_sub, count = shard_range(R, rank, world)
    part = np.zeros((n_leaves, 4), dtype=np.int64)
    for i, o in enumerate(leaves):
        for r in range(first_sub, first_sub + count):
            st, pts, _ = o.leaf_rollout(seed, 500 + i, r, 1, True)
            if st == 0:
                part[i] += pts
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)                                   # integer sum: order independent, bit reproducible
    if rank == 0:
        full = np.zeros((n_leaves, 4), dtype=np.int64)
        for i, o in enumerate(leaves):
            for r in range(R):
                st, pts, _ = o.leaf_rollout(seed, 500 + i, r, 1, True)
                if st == 0:
                    full[i] += pts
        ret["ok"] = bool(np.array_equal(t.numpy(), full))
    dist.destroy_process_group()


def test_root_stats_sum_over_two_ranks_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True


def _pimc_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostsim_lib
    import oracle_lib
    from oracle_lib import Fdo

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, sim = oracle_lib.load(), hostsim_lib.load()
    seed, n_roots, n_det, R = 7, 5, 9, 8
    roots = []
    for i in range(n_roots):
        o = Fdo.new_game_philox(L, seed, i, 0)
        for k in range(6 + 7 * i):
            o.random_step(seed, i, 0, True, k)
        roots.append(o)
    first_sub, count = shard_range(n_det, rank, world)
    stats = np.zeros((n_roots, 80), dtype=np.int64)
    for i, o in enumerate(roots):
        rows = np.zeros((count, 39), dtype=np.uint32)
        status = np.zeros(count, dtype=np.uint8)
        for d in range(count):
            st, vis, _ = o.flat_mc(seed, 100 + i, first_sub + d, R, 2)
            rows[d], status[d] = vis, st
        if count:
            sim.sim_root_stats(hostsim_lib.ptr(rows), hostsim_lib.ptr(status), count, o.allowed(), hostsim_lib.ptr(stats[i]))
    t = torch.from_numpy(stats.copy())
    dist.all_reduce(t)                                   # the exchange step: integer sum of [roots][80] statistics
    total = t.numpy()
    if rank == 0:
        ok = True
        for i, o in enumerate(roots):
            rows = [o.flat_mc(seed, 100 + i, d, R, 2) for d in range(n_det)]
            good = np.array([v for st, v, _ in rows if st == 0])
            for strategy in (0, 1):                      # R = 8: every row total is a power of two, so Average agrees exactly as well
                pick = sim.sim_root_pick(strategy, hostsim_lib.ptr(np.ascontiguousarray(total[i])), o.allowed())
                ok = ok and pick == oracle_lib.fuse(L, strategy, good, o.allowed())
        ret["ok"] = bool(ok)
    dist.destroy_process_group()


def test_pimc_decision_over_two_ranks_gloo():
    """Each rank evaluates its share of the determinizations (oracle rows as a stand-in for the kernel), reduces them to the additive
    root statistics, gloo all-reduces them, and the pick equals the PolicyFusion decision over all determinizations."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pimc_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
