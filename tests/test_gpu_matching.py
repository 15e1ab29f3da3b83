"""GPU parity of the determinizer and the fused determinize→rollout kernel (configs 3 and 4) against the oracle."""
import numpy as np
import pytest

from oracle_lib import DK_STATE_DTYPE, Fdo

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def midgame_states(orc, n, first, no_solo_every=2):
    """Oracle games advanced to card_index in {8,16,24,32} round-robin (BASELINE config 3), some forced to Normal/Wedding."""
    prng = np.random.default_rng(5)
    objs = []
    for i in range(n):
        o = Fdo.new_game_philox(orc, SEED, first + i, 3)
        target = (8, 16, 24, 32)[i % 4]
        while o.allowed() and o.info()["card_index"] < target:
            m = o.allowed()
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if i % no_solo_every == 0 and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 else 24
            o.play(a)
        objs.append(o)
    return objs


def to_dev(objs):
    import torch

    recs = np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE)
    return torch.from_numpy(np.frombuffer(recs.tobytes(), dtype=np.uint8).reshape(len(objs), 128).copy()).cuda()


def test_determinize_matches_oracle_and_is_consistent(dk, orc):
    import torch

    n, S = 96, 40
    objs = midgame_states(orc, n, 0)
    hands, res, status = dk.determinize(1, to_dev(objs), S, dk.rng(SEED, 10_000, 1))
    torch.cuda.synchronize()
    hands, res, status = hands.cpu().numpy().astype(np.uint64), res.cpu().numpy(), status.cpu().numpy()
    for i, o in enumerate(objs):
        for s in range(S):
            st_o, h_o, r_o = o.card_matching(SEED, 10_000 + i, s, 1)
            assert (int(status[i, s]), [int(x) for x in hands[i, s]], list(res[i, s])) == (st_o, h_o, r_o), f"state {i} sample {s}"
            if st_o == 0:
                assert o.is_consistent(h_o, r_o) == 0


def test_determinize_many_samples_all_consistent(dk, orc):
    """4096 samples per info-state (config 3's per-state count) on a few states: every sample passes is_consistent."""
    import torch

    objs = midgame_states(orc, 8, 500)
    S = 4096
    hands, res, status = dk.determinize(1, to_dev(objs), S, dk.rng(SEED, 0, 2))
    torch.cuda.synchronize()
    hands, res, status = hands.cpu().numpy().astype(np.uint64), res.cpu().numpy(), status.cpu().numpy()
    assert int(status.max()) == 0
    for i, o in enumerate(objs):
        for s in range(0, S, 7):
            assert o.is_consistent([int(x) for x in hands[i, s]], list(res[i, s])) == 0
        # the observer keeps its real hand in every sample
        obs = o.info()["current_player"]
        assert (hands[i, :, obs] == np.uint64(o.hands()[obs])).all()


@pytest.mark.parametrize("determinize", [True, False])
def test_leaf_rollouts_sum_matches_oracle(dk, orc, determinize):
    import torch

    n, R = 24, 300
    objs = midgame_states(orc, n, 900)
    sums = dk.leaf_rollouts(to_dev(objs), R, dk.rng(SEED, 7000, 4), determinize=determinize)
    torch.cuda.synchronize()
    sums = sums.cpu().numpy()
    for i, o in enumerate(objs):
        exp = np.zeros(4, dtype=np.int64)
        for r in range(R):
            st, pts, _ = o.leaf_rollout(SEED, 7000 + i, r, 4, determinize)
            if st == 0:
                exp += np.array(pts)
        assert np.array_equal(sums[i], exp), f"leaf {i}"


def test_doko_assignment_matches_oracle(dk, orc):
    """rs-doko-assignment on the GPU (config 3, simplified engine) == the Vec-based oracle restatement."""
    import torch

    from oracle_lib import Doko

    prng = np.random.default_rng(17)
    objs = []
    for i in range(64):
        o = Doko.new_game_philox(orc, SEED, 300 + i, 0)
        target = 4 + (8, 16, 24, 32)[i % 4]
        n = 0
        while o.allowed() and n < target:
            m = o.allowed()
            legal = [a for a in range(26) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if i % 2 == 0 and (m >> 25) & 1:
                a = 25
            o.play(a)
            n += 1
        objs.append(o)
    S = 48
    hands, res, status = dk.determinize(0, to_dev(objs), S, dk.rng(SEED, 5000, 6))
    torch.cuda.synchronize()
    hands, status = hands.cpu().numpy().astype(np.uint64), status.cpu().numpy()
    for i, o in enumerate(objs):
        for s in range(S):
            st_o, h_o = o.sample_assignment(SEED, 5000 + i, s, 6)
            assert (int(status[i, s]), [int(x) for x in hands[i, s]]) == (st_o, h_o), f"state {i} sample {s}"


def test_doko_assignment_support_sets_on_gpu(dk):
    """The reference's own rs-doko-assignment tests on the kernel: 1000 samples → exactly the enumerated 9 / 12 assignments."""
    import torch

    from assignment_cases import CASES, canonical, case_record, expected_set

    for case in CASES:
        rec = case_record(case)
        st = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).reshape(1, 128).copy()).cuda()
        hands, _, status = dk.determinize(0, st, 1000, dk.rng(42, 0, 0))
        torch.cuda.synchronize()
        assert int(status.max()) == 0
        seen = {canonical([int(x) for x in h]) for h in hands[0].cpu().numpy().astype(np.uint64)}
        assert len(seen) == case["n_unique"] and expected_set(case) <= seen


def test_card_matching_support_set_on_gpu(dk, orc):
    """10^5 GPU samples per late-game info-state (<= 10 hidden cards) all land inside the enumerated set of is_consistent assignments
    (tests/test_card_matching_support.py builds the same sets for the oracle), and together they reach practically all of it: the greedy
    rules of card_matching leave no consistent assignment of these states unreachable."""
    import torch

    from oracle_lib import Bulk

    b = Bulk(orc, 1, 96, SEED, first_id=4000, epoch=3, mode=2)
    S = 100_000
    hands, _, status = dk.determinize(1, to_dev_bytes(b.bytes()), S, dk.rng(SEED, 4000, 8))
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    hands = hands.cpu().numpy().astype(np.uint64)
    total = reached = 0
    key = [("h", "<u8", 4)]
    for i in range(b.n):
        support = Fdo.from_dk_state(orc, b.recs[i:i + 1]).consistent_hands()
        seen = np.unique(np.ascontiguousarray(hands[i]).view(key))
        assert np.isin(seen, np.ascontiguousarray(support).view(key)).all(), f"state {i}: a GPU sample outside the consistent set"
        total += len(support)
        reached += len(seen)
    assert total > 5000 and reached >= 0.999 * total, (reached, total)


def to_dev_bytes(raw):
    import torch

    return torch.from_numpy(raw).cuda()
