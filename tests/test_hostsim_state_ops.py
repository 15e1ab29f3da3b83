"""Device logic (compiled for the CPU by tests/hostsim) vs the oracle, state by state.

For seeded random games of both engines every step checks: the 128-byte dk_state record, the legal mask, the encoders
(311 / 110 / 114 tokens) and a resume-playout from that state.  The same comparisons run against the real kernels in
tests/test_gpu_state_ops.py (-m gpu).
"""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Doko, Fdo

SEED = 0x5EEDD0C0


def new_rec(sim, hands, start):
    rec = np.zeros(1, dtype=DK_STATE_DTYPE)
    sim.sim_new_game(hostsim_lib.ptr(rec), (C.c_uint64 * 4)(*hands), start)
    return rec


def sim_encode(sim, layout, rec, n):
    out = np.zeros(n, dtype=np.int64)
    sim.sim_encode(layout, hostsim_lib.ptr(rec), hostsim_lib.ptr(out))
    return out


def sim_playout(sim, engine, rec, seed, unit, unit_hi, epoch, with_ann):
    pts = np.zeros(4, dtype=np.int32)
    steps = C.c_uint32()
    sim.sim_playout_from_state(engine, hostsim_lib.ptr(rec), seed, unit, unit_hi, epoch, int(with_ann), hostsim_lib.ptr(pts), C.byref(steps))
    return list(pts), steps.value


def forced_reservations(kind, g):
    """Reservation policies that reach the rare game types often (random play is ~99.97 % solos)."""
    return {"random": None, "healthy": 24, "wedding_if_possible": 25}[kind]


@pytest.mark.parametrize("kind", ["random", "healthy", "wedding_if_possible"])
def test_fdo_state_walk(orc, kind):
    sim = hostsim_lib.load()
    rng = np.random.default_rng(1234)
    n_games = 60 if kind == "random" else 120
    for g in range(n_games):
        o = Fdo.new_game_philox(orc, SEED, g, 7)
        rec = new_rec(sim, o.hands(), o.info()["current_player"])
        step = 0
        while True:
            assert rec.tobytes() == np.array([o.export()], dtype=DK_STATE_DTYPE).tobytes(), f"game {g} step {step}: record differs"
            legal = int(sim.sim_legal_mask(1, hostsim_lib.ptr(rec)))
            assert legal == o.allowed()
            assert np.array_equal(sim_encode(sim, 2, rec, 311), o.encode_pi())
            if step % 7 == 0:
                for wa in (False, True):
                    assert sim_playout(sim, 1, rec, SEED, g, step, 3, wa) == tuple(o.rollout(SEED, g, step, 3, wa)), f"game {g} step {step} playout"
            if legal == 0:
                break
            acts = [a for a in range(39) if (legal >> a) & 1]
            a = int(rng.choice(acts))
            forced = forced_reservations(kind, g)
            if forced is not None and o.info()["phase"] == 0:
                a = forced if (legal >> forced) & 1 else 24
            assert sim.sim_apply(1, hostsim_lib.ptr(rec), a, 0) == 0
            o.play(a)
            step += 1
        assert step >= 52


def test_fdo_illegal_action_leaves_state_unchanged(orc):
    sim = hostsim_lib.load()
    o = Fdo.new_game_philox(orc, SEED, 1, 0)
    rec = new_rec(sim, o.hands(), o.info()["current_player"])
    before = rec.tobytes()
    for a in (0, 33, 38, 39, 200):
        assert sim.sim_apply(1, hostsim_lib.ptr(rec), a, 0) == 1
        assert rec.tobytes() == before


def test_fdo_skip_single_matches_az_env(orc):
    """take_action_by_action_index(.., skip_single=true, ..): auto-play while exactly one non-call action is legal."""
    sim = hostsim_lib.load()
    rng = np.random.default_rng(5)
    calls = 0x1F << 33
    for g in range(40):
        o = Fdo.new_game_philox(orc, SEED, 1000 + g, 0)
        rec = new_rec(sim, o.hands(), o.info()["current_player"])
        while o.allowed():
            legal = o.allowed()
            a = int(rng.choice([x for x in range(39) if (legal >> x) & 1]))
            assert sim.sim_apply(1, hostsim_lib.ptr(rec), a, 1) == 0
            o.play(a)
            while o.allowed() and bin(o.allowed() & ~calls).count("1") == 1:
                o.play((o.allowed() & ~calls).bit_length() - 1)
            assert rec.tobytes() == np.array([o.export()], dtype=DK_STATE_DTYPE).tobytes()


def test_doko_state_walk(orc):
    sim = hostsim_lib.load()
    rng = np.random.default_rng(99)
    for g in range(150):
        o = Doko.new_game_philox(orc, SEED, g, 1)
        rec = new_rec(sim, o.hands(), o.info()["current_player"])
        step = 0
        while True:
            assert rec.tobytes() == np.array([o.export()], dtype=DK_STATE_DTYPE).tobytes(), f"game {g} step {step}: record differs"
            legal = int(sim.sim_legal_mask(0, hostsim_lib.ptr(rec)))
            assert legal == o.allowed()
            assert np.array_equal(sim_encode(sim, 0, rec, 110), o.encode(False))
            assert np.array_equal(sim_encode(sim, 1, rec, 114), o.encode(True))
            if legal == 0:
                break
            if step % 9 == 0:
                c = Doko(orc, orc.orc_doko_clone(o.h))
                while c.random_step(SEED, g, 5) >= 0:
                    pass
                pts, steps = sim_playout(sim, 0, rec, SEED, g, 0, 5, False)
                assert pts == c.info()["points"] and steps == c.info()["n_play_actions"] - o.info()["n_play_actions"]
            acts = [a for a in range(26) if (legal >> a) & 1]
            a = int(rng.choice(acts))
            if g % 2 == 0 and (legal >> 25) & 1:
                a = 25
            assert sim.sim_apply(0, hostsim_lib.ptr(rec), a, 0) == 0
            o.play(a)
            step += 1
        assert step == 52
