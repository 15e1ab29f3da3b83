"""GPU parity of the lock-step self-play driver through the C ABI against oracle/selfplay.hpp: every experience
row (311 tokens, policy target, value target, mover) bit for bit, in the driver's deterministic row order."""
import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def run_driver(dk, n, first_id, az_epoch, keep_prob, capacity, flags=0, max_turns=260):
    import torch

    states = dk.new_games(1, n, dk.rng(SEED, first_id, 0))
    sp = dk.self_play(n, capacity)
    sp.reset()
    turns = 0
    for t in range(max_turns):
        sp.begin_turn(states, az_epoch, keep_prob, dk.rng(SEED, first_id, t), flags)
        _, fl, _ = sp.turn_view()
        if t % 8 == 0 and bool(((fl & 1) != 0).all()):
            break
        sp.uniform_search(dk.rng(SEED, first_id, t))
        err = sp.end_turn(states)
        turns += 1
    assert int(err.max()) == 0
    sp.finalize(states)
    rows, dropped, unfinished = sp.counts()
    torch.cuda.synchronize()
    return sp, states, rows, dropped, unfinished


@pytest.mark.parametrize("n,az_epoch,keep_prob", [(300, 0, 1.0), (257, 12, 0.5), (64, 3, 0.0)])
def test_driver_matches_oracle(dk, orc, n, az_epoch, keep_prob):
    first_id = 9_000
    sp, states, rows, dropped, unfinished = run_driver(dk, n, first_id, az_epoch, keep_prob, capacity=n * 130)
    assert dropped == 0 and unfinished == 0
    games = [oracle_lib.selfplay_uniform(orc, SEED, first_id + g, az_epoch, keep_prob) for g in range(n)]
    # expected order: turn-major, game order inside a turn
    order = sorted((int(t), g, i) for g, gm in enumerate(games) for i, t in enumerate(gm["turn"]))
    assert rows == len(order)
    st = sp.states[:rows].cpu().numpy()
    po = sp.policy[:rows].cpu().numpy()
    va = sp.value[:rows].cpu().numpy()
    pl = sp.player[:rows].cpu().numpy()
    ga = sp.game[:rows].cpu().numpy()
    for r, (t, g, i) in enumerate(order):
        gm = games[g]
        assert ga[r] == g, (r, t, g)
        assert pl[r] == gm["player"][i]
        assert (st[r] == gm["states"][i]).all(), (r, t, g)
        assert (po[r].view(np.uint32) == gm["policy"][i].view(np.uint32)).all(), (r, t, g, po[r], gm["policy"][i])
        assert (va[r].view(np.uint32) == gm["value"][i].view(np.uint32)).all(), (r, t, g)
    # final points of every game
    pts = states.cpu().numpy()[:, 120:124].view(np.int8)
    for g, gm in enumerate(games):
        if len(gm["turn"]):
            assert list(pts[g]) == gm["points"]


def test_capacity_overflow_drops_rows_and_keeps_playing(dk, orc):
    n = 200
    sp, states, rows, dropped, unfinished = run_driver(dk, n, 77, 0, 1.0, capacity=5000)
    total = sum(oracle_lib.selfplay_uniform(orc, SEED, 77 + g, 0, 1.0)["turns"] for g in range(n))
    assert rows == 5000 and dropped == total - 5000 and unfinished == 0
    # the first 5000 rows are the same rows an unbounded buffer would hold
    sp2, _, rows2, _, _ = run_driver(dk, n, 77, 0, 1.0, capacity=n * 130)
    assert rows2 == total
    assert bool((sp.states[:5000] == sp2.states[:5000]).all()) and bool((sp.value[:5000] == sp2.value[:5000]).all())


def test_search_forced_flag_records_every_turn(dk):
    import master_doko_reinforcement_learning_b200.api as api

    n = 128
    sp, states, rows, dropped, unfinished = run_driver(dk, n, 5, 0, 0.0, capacity=n * 130, flags=api.SP_SEARCH_FORCED)
    sp0, _, rows0, _, _ = run_driver(dk, n, 5, 0, 1.0, capacity=n * 130)
    assert rows == rows0 and dropped == 0          # keep_prob is irrelevant when forced moves are searched like any other
    assert bool((sp.states[:rows] == sp0.states[:rows]).all())


def test_illegal_action_is_reported_and_state_unchanged(dk):
    import torch

    n = 64
    states = dk.new_games(1, n, dk.rng(SEED, 0, 0))
    before = states.clone()
    sp = dk.self_play(n, n * 4)
    sp.begin_turn(states, 0, 1.0, dk.rng(SEED, 0, 0), 1)
    policy = torch.zeros((n, 39), dtype=torch.float32, device="cuda")
    action = torch.full((n,), 5, dtype=torch.uint8, device="cuda")      # a card in the reservation phase
    err = sp.end_turn(states, policy, action)
    assert int(err.min()) == 1 and bool((states == before).all())
