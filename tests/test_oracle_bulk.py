"""The bulk oracle checkers (oracle/capi.cpp orc_bulk_*) against the one-at-a-time oracle functions they batch — so that the full-size
GPU parity tests compare against the SAME restatement that the golden-vector tests pin."""
import numpy as np

from oracle_lib import Bulk, Doko, Fdo

SEED = 0xD0C05EED


def test_bulk_card_matching_equals_single_calls(orc):
    b = Bulk(orc, 1, 40, SEED, first_id=100, epoch=3, mode=1)
    hands, res, status, cons, _ = b.determinize(6, epoch=5, first_sub=2)
    phases = set()
    for i in range(b.n):
        o = Fdo.from_dk_state(orc, b.recs[i:i + 1])
        phases.add(o.info()["phase"])
        for s in range(6):
            st, h, r = o.card_matching(SEED, 100 + i, 2 + s, 5)
            assert (st, h, r) == (int(status[i, s]), [int(x) for x in hands[i, s]], list(res[i, s]))
            assert cons[i, s] == (o.is_consistent(h, r) if st == 0 else -1)
    assert {0, 1, 2} <= phases          # reservation-, announcement- and card-phase info-states are all in the recipe
    assert int(status.max()) == 0 and int(np.abs(cons).max()) == 0


def test_bulk_sample_assignment_equals_single_calls(orc):
    b = Bulk(orc, 0, 24, SEED, first_id=7, epoch=0, mode=0)
    hands, _, status, _, _ = b.determinize(5, epoch=6)
    assert int(status.max()) == 0
    for i in range(b.n):
        o = Doko.new_game_philox(orc, SEED, 7 + i, 0)                   # the recipe: deal, then 4 + 8 / 16 / 24 / 32 random actions
        for _ in range(4 + 8 * (1 + (i & 3))):
            o.random_step(SEED, 7 + i, 0)
        assert o.export().tobytes() == b.recs[i].tobytes()
        for s in range(5):
            st, h = o.sample_assignment(SEED, 7 + i, s, 6)
            assert (st, h) == (int(status[i, s]), [int(x) for x in hands[i, s]])


def test_bulk_leaf_rollouts_equal_single_calls(orc):
    b = Bulk(orc, 1, 10, SEED, first_id=900, epoch=3, mode=0)
    for det in (True, False):
        sums, _ = b.leaf_rollouts(20, epoch=4, determinize=det)
        for i in range(b.n):
            o = Fdo.from_dk_state(orc, b.recs[i:i + 1])
            exp = np.zeros(4, dtype=np.int64)
            for r in range(20):
                st, pts, _ = o.leaf_rollout(SEED, 900 + i, r, 4, det)
                if st == 0:
                    exp += np.array(pts)
            assert np.array_equal(sums[i], exp)


def test_bulk_step_equals_single_calls(orc):
    b = Bulk(orc, 1, 12, SEED, first_id=40, epoch=3, mode=1)
    objs = [Fdo.from_dk_state(orc, b.recs[i:i + 1]) for i in range(b.n)]
    for k in range(6):
        act, recs, obs = b.step(60 + k, want_recs=True, want_obs=True)
        for i, o in enumerate(objs):
            a = o.step_site(SEED, 40 + i, 60 + k, True, False)
            assert a == (int(act[i]) if act[i] != 0xFF else -1)
            assert o.export().tobytes() == recs[i].tobytes()
            assert np.array_equal(o.encode_pi(), obs[i])
