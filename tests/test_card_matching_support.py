"""Support set of the full-rules determinizer, the counterpart of rs-doko-assignment's own support-set tests
(assignment.rs:865-1109) for card_matching, whose reference tests are OS-seeded: for late-game info-states (<= 10 hidden cards) EVERY
assignment that passes is_consistent is enumerated, and the sampler's draws must all land inside that set.  The members the greedy rules
(forced card → forced seat → forced ♣Q → random card to the FIRST seat that can hold it, card_matching.rs:78-204) can never reach are
counted: rule 4's "first eligible seat" makes the sampler non-uniform and incomplete by design, and the test records by how much."""
import numpy as np

from oracle_lib import Bulk, Fdo

SEED = 0xD0C05EED


def support_of(orc, rec):
    o = Fdo.from_dk_state(orc, rec)
    members = o.consistent_hands()
    return o, {tuple(int(x) for x in m) for m in members}


def test_oracle_card_matching_samples_stay_inside_the_enumerated_support(orc):
    b = Bulk(orc, 1, 140, SEED, first_id=4000, epoch=3, mode=2)
    S = 1500
    hands, res, status, cons, _ = b.determinize(S, epoch=8)
    assert int(status.max()) == 0 and int(np.abs(cons).max()) == 0
    reached_frac, sizes = [], []
    for i in range(b.n):
        o, support = support_of(orc, b.recs[i:i + 1])
        real = tuple(int(x) for x in b.recs[i]["hands"])
        assert real in support                                       # the true deal is always a member
        seen = {tuple(int(x) for x in h) for h in hands[i]}
        assert seen <= support, f"state {i}: a sample outside the consistent set"
        sizes.append(len(support))
        reached_frac.append(len(seen) / len(support))
    # the states are not trivial, and the sampler covers a good part of the support (a reachability regression would show up here)
    assert max(sizes) >= 300 and np.mean([s > 1 for s in sizes]) > 0.6
    assert np.mean(reached_frac) > 0.5
