"""Self-play driver on the CPU: properties of the oracle's self_play restatement (oracle/selfplay.hpp) that the
reference's loop guarantees (self_play.rs:56-207), and the device helper functions against it."""
import numpy as np

import hostsim_lib
import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

SEED = 0x5E1F


def test_self_play_rows_have_the_reference_shape(orc):
    for unit in range(12):
        for az_epoch, keep in ((0, 1.0), (0, 0.0), (12, 0.4)):
            g = oracle_lib.selfplay_uniform(orc, SEED, unit, az_epoch, keep)
            n = len(g["turn"])
            assert n <= g["turns"] <= 250
            assert (np.diff(g["turn"].astype(int)) > 0).all()
            forced = g["forced"].astype(bool)
            if keep == 1.0:
                assert n == g["turns"]                              # every turn is recorded
            if keep == 0.0:
                assert not forced.any()                             # forced moves are never kept, searched moves always
            # policy targets: one-hot for forced moves, a distribution over the allowed actions otherwise
            assert np.allclose(g["policy"].sum(axis=1), 1.0, atol=1e-6)
            assert ((g["policy"][forced] == 1.0).sum(axis=1) == 1).all()
            if az_epoch < 10:
                assert (g["policy"][:, 33:38] == 0).all()           # calls are masked below MIN_EPOCH (full_doko.rs:23,84-90)
            # value targets: final rewards (points / 8) rotated to the row's mover
            pts = np.array(g["points"], dtype=np.float32) / 8.0
            for r in range(n):
                assert (g["value"][r] == np.roll(pts, -int(g["player"][r]))).all()
            assert sum(g["points"]) == 0
            # the phase token of every recorded observation is a non-terminal phase
            assert (g["states"][:, 310] != 3).all() if n else True


def test_keep_probability_only_thins_forced_moves(orc):
    full = oracle_lib.selfplay_uniform(orc, SEED, 3, 0, 1.0)
    thin = oracle_lib.selfplay_uniform(orc, SEED, 3, 0, 0.5)
    assert full["turns"] == thin["turns"]                          # same game: the keep draw does not touch the game stream
    kept = set(thin["turn"].tolist())
    for t, f in zip(full["turn"].tolist(), full["forced"].tolist()):
        if not f:
            assert t in kept
    assert 0 < len(kept) < len(full["turn"])
    rows = {int(t): i for i, t in enumerate(full["turn"])}
    for i, t in enumerate(thin["turn"].tolist()):
        assert (thin["states"][i] == full["states"][rows[t]]).all()


def test_device_helpers_match_oracle(orc):
    sim = hostsim_lib.load()
    prng = np.random.default_rng(2)
    for w in [0, 1, 255, 256, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF] + prng.integers(0, 1 << 32, size=200).tolist():
        assert sim.sim_sp_keep_draw(int(w)) == np.float32((int(w) >> 8) * 2.0 ** -24)
    n_states = 0
    for g in range(10):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        while o.allowed():
            rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
            for az_epoch in (0, 9, 10, 500):
                assert sim.sim_sp_az_allowed(hostsim_lib.ptr(rec), az_epoch) == orc.orc_fdo_az_allowed(o.h, 0, az_epoch)
            m = o.allowed()
            legal = [a for a in range(39) if (m >> a) & 1]
            o.play(int(prng.choice(legal)))
            n_states += 1
        rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
        pts = rec["points"][0].astype(np.float32) / 8.0
        for player in range(4):
            for k in range(4):
                assert sim.sim_sp_value_target(hostsim_lib.ptr(rec), player, k) == pts[(player + k) % 4]
    assert n_states > 500
