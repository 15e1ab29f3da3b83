"""GPU parity of the PIMC move decision through the C ABI: the flat Monte-Carlo evaluator against oracle/pimc.hpp
(visit counts and exact value sums per determinization), the PolicyFusion kernels against the oracle restatement and the reference's
own four known answers, and the sharded statistics path against the single-shot fuse."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def states_at_random_depth(orc, n, first, max_depth=80):
    prng = np.random.default_rng(first)
    objs = []
    for i in range(n):
        o = Fdo.new_game_philox(orc, SEED, first + i, 4)
        for _ in range(int(prng.integers(0, max_depth))):
            m = o.allowed()
            if not m:
                break
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if i % 2 == 0 and (m >> 24) & 1:                      # half of the games stay Normal / Wedding
                a = 25 if (m >> 25) & 1 else 24
            o.play(a)
        objs.append(o)
    return objs


def to_dev(objs):
    import torch

    recs = np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE)
    return torch.from_numpy(np.frombuffer(recs.tobytes(), dtype=np.uint8).reshape(len(objs), 128).copy()).cuda()


def mask_of(actions):
    m = 0
    for a in actions:
        m |= 1 << a
    return m


# (n_roots, n_det, n_rollouts, first_sub): several determinizations per block / ragged last block / one determinization per block
# with several passes over the rollouts / a non-power-of-two rollout count / an offset determinization range (rank > 0)
@pytest.mark.parametrize("n,n_det,R,first_sub", [(40, 5, 8, 0), (24, 37, 3, 0), (6, 2, 200, 0), (16, 6, 16, 11)])
def test_flat_mc_matches_oracle(dk, orc, n, n_det, R, first_sub):
    import torch

    objs = states_at_random_depth(orc, n, 1000 * n_det + R)
    first_id = 77_000
    visits, values, status = dk.pimc_evaluate(to_dev(objs), n_det, R, dk.rng(SEED, first_id, 5, first_sub))
    torch.cuda.synchronize()
    visits, values, status = visits.cpu().numpy().view(np.uint32), values.cpu().numpy(), status.cpu().numpy()
    phases = set()
    for i, o in enumerate(objs):
        phases.add(o.info()["phase"])
        for d in range(n_det):
            st_o, vis_o, val_o = o.flat_mc(SEED, first_id + i, first_sub + d, R, 5)
            assert int(status[i, d]) == st_o, (i, d)
            assert (visits[i, d] == vis_o).all(), (i, d, visits[i, d], vis_o)
            assert (values[i, d] == val_o).all(), (i, d, values[i, d], val_o)
            if st_o == 0 and o.allowed():
                assert int(visits[i, d].sum()) == R
    assert len(phases) >= 3


def test_decision_end_to_end(dk, orc):
    """evaluate → legal mask → fuse on the device == oracle flat_mc rows → oracle fuse, for both strategies; the sharded path (two
    'ranks' with half of the determinizations each → root statistics → sum → pick) gives the same MaxN decision."""
    import torch

    n, n_det, R = 64, 16, 32
    objs = states_at_random_depth(orc, n, 4242, max_depth=60)
    dev = to_dev(objs)
    first_id = 5_000
    visits, _, status = dk.pimc_evaluate(dev, n_det, R, dk.rng(SEED, first_id, 6), want_values=False)
    allowed = dk.legal_mask(1, dev)
    act = {s: dk.fuse(s, visits, allowed, status) for s in (0, 1)}
    half = n_det // 2
    stats = None
    for rank in range(2):
        v_r, _, st_r = dk.pimc_evaluate(dev, half, R, dk.rng(SEED, first_id, 6, rank * half), want_values=False)
        part = dk.pimc_root_stats(v_r, allowed, st_r)
        stats = part if stats is None else stats + part
    pick = {s: dk.pimc_pick(s, stats, allowed) for s in (0, 1)}
    torch.cuda.synchronize()
    allowed_h = allowed.cpu().numpy().astype(np.uint64)
    for i, o in enumerate(objs):
        rows = []
        for d in range(n_det):
            st_o, vis_o, _ = o.flat_mc(SEED, first_id + i, d, R, 6)
            if st_o == 0:
                rows.append(vis_o)
        assert int(allowed_h[i]) == o.allowed()
        for s in (0, 1):
            a_dev, n_ok = int(act[s][0][i]), int(act[s][1][i])
            assert n_ok == len(rows)
            if not o.allowed():
                continue
            want = oracle_lib.fuse(orc, s, np.array(rows), o.allowed())
            assert a_dev == want, (i, s)
            assert (o.allowed() >> a_dev) & 1
            assert int(pick[s][i]) == want, (i, s)                # R = 32: the f32 sums are exact, so Average agrees as well


def test_reference_fusion_known_answers(dk):
    import torch

    cases = json.load(open(os.path.join(HERE, "golden", "policy_fusion_cases.json")))["cases"]
    for c in cases:
        v = torch.tensor(np.array(c["visits"], dtype=np.int32)).cuda().reshape(1, -1, 39)
        allowed = torch.tensor([mask_of(c["allowed"])], dtype=torch.int64).cuda()
        strategy = 0 if c["strategy"] == "PolicyFusionMaxN" else 1
        action, n_ok = dk.fuse(strategy, v, allowed)
        assert int(action[0]) == c["expected"] and int(n_ok[0]) == len(c["visits"]), c["name"]


def test_fuse_matches_oracle_on_random_rows(dk, orc):
    """Ties, stray visits on not-allowed actions, all-zero rows (NaN path of the Average strategy), failed samples, roots without
    any successful sample."""
    import torch

    prng = np.random.default_rng(17)
    n, rows = 600, 23
    v = np.zeros((n, rows, 39), dtype=np.uint32)
    allowed = np.zeros(n, dtype=np.uint64)
    status = (prng.random((n, rows)) < 0.2).astype(np.uint8)
    status[::50] = 1
    for i in range(n):
        legal = prng.choice(39, size=int(prng.integers(1, 13)), replace=False)
        allowed[i] = mask_of(legal.tolist())
        kind = i % 3
        if kind == 0:
            v[i][:, legal] = prng.integers(0, 3, size=(rows, len(legal)))
        elif kind == 1:
            v[i] = prng.integers(0, 50, size=(rows, 39))
            v[i, ::5] = 0
        else:
            v[i][:, legal] = prng.integers(0, 1 << 31, size=(rows, len(legal)))
    vd = torch.from_numpy(v.view(np.int32)).cuda()
    ad = torch.from_numpy(allowed.view(np.int64)).cuda()
    sd = torch.from_numpy(status).cuda()
    for strategy in (0, 1):
        action, n_ok = dk.fuse(strategy, vd, ad, sd)
        action, n_ok = action.cpu().numpy(), n_ok.cpu().numpy()
        for i in range(n):
            ok = v[i][status[i] == 0]
            assert n_ok[i] == len(ok)
            if len(ok) == 0:
                assert action[i] == 0xFF
            else:
                assert action[i] == oracle_lib.fuse(orc, strategy, ok, int(allowed[i])), (i, strategy)


def test_pimc_decide_falls_back_to_a_random_action_without_successful_samples(dk, orc):
    """DefaultImpiPolicy::execute (compare_impi.rs:357-368): when no determinization succeeded, a random allowed non-announcement action
    is played.  Forced here by marking every sample of some roots as failed; the other roots keep the fused decision."""
    import torch

    from master_doko_reinforcement_learning_b200.sharding import pimc_decide
    from oracle_lib import Bulk

    b = Bulk(orc, 1, 64, SEED, first_id=300, epoch=3, mode=1)
    states = torch.from_numpy(b.bytes()).cuda()
    real = dk.pimc_evaluate

    def failing(states_, n_det, n_rollouts, rng, want_values=True, stream=None):
        v, val, st = real(states_, n_det, n_rollouts, rng, want_values=want_values, stream=stream)
        st[::4] = 1                                                   # every sample of every fourth root is a dead end
        return v, val, st

    dk.pimc_evaluate = failing
    try:
        act, stats = pimc_decide(dk, states, 8, 0, SEED, first_id=300, epoch=5, n_rollouts=8)
    finally:
        dk.pimc_evaluate = real
    torch.cuda.synchronize()
    allowed = dk.legal_mask(1, states)
    v, _, st = dk.pimc_evaluate(states, 8, 8, dk.rng(SEED, 300, 5), want_values=False)
    fused = dk.fuse(0, v, allowed, st)[0].cpu().numpy()
    fallback = dk.random_action(1, states, dk.rng(SEED, 300, 5), flags=0).cpu().numpy()
    act = act.cpu().numpy()
    assert (stats[::4, 78] == 0).all()                                # no successful sample counted for the forced roots
    assert np.array_equal(act[::4], fallback[::4])
    keep = np.ones(64, dtype=bool); keep[::4] = False
    assert np.array_equal(act[keep], np.where(fused[keep] == 0xFF, fallback[keep], fused[keep]))
    legal = allowed.cpu().numpy().astype(np.uint64)
    calls = np.uint64(0x1F << 33)
    for i in range(0, 64, 4):                                         # the fallback is an allowed non-announcement action
        assert (int(legal[i]) & ~int(calls)) >> int(act[i]) & 1
