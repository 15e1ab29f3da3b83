"""Determinizer device logic (hostsim) vs the oracle's card_matching restatement, plus the reference's own property:
every sample must pass is_consistent (rs-full-doko/src/matching/card_matching.rs:548-640, rs-full-doko-cmd/src/main.rs:190-283)."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

SEED = 1711


def sim_determinize(sim, rec, unit, sample, epoch=0):
    hands = (C.c_uint64 * 4)()
    res = (C.c_uint8 * 4)()
    st = sim.sim_fdo_determinize(hostsim_lib.ptr(rec), SEED, unit, sample, epoch, hands, res)
    return st, [int(x) for x in hands], list(res)


@pytest.mark.parametrize("kind", ["random", "no_solo"])
def test_card_matching_soak(orc, kind):
    """Seeded games; at EVERY state several samples: device == oracle bit for bit, and every sample is consistent."""
    sim = hostsim_lib.load()
    prng = np.random.default_rng(11)
    n_states = n_dead = 0
    for g in range(40):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        step = 0
        while o.allowed():
            rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
            for sample in range(3):
                st_o, hands_o, res_o = o.card_matching(SEED, g * 1000 + step, sample)
                st_d, hands_d, res_d = sim_determinize(sim, rec, g * 1000 + step, sample)
                assert (st_d, hands_d, res_d) == (st_o, hands_o, res_o), f"game {g} step {step} sample {sample}"
                if st_o == 0:
                    assert o.is_consistent(hands_o, res_o) == 0, f"game {g} step {step} sample {sample}: inconsistent sample"
                else:
                    n_dead += 1
                n_states += 1
            # every state of the game, so leaves INSIDE a trick are covered: the chained card draws of the running trick continue from
            # the legal-card counts of the plays already made (bridge: FdoResume::chain_mul; determinized: fdo_live_with_sample, which
            # the simulator entry point checks against bridging the determinized record)
            for det in (1, 0):
                pts = np.zeros(4, dtype=np.int32)
                steps = C.c_uint32()
                st = sim.sim_fdo_leaf_rollout(hostsim_lib.ptr(rec), SEED, g, step, 9, det, hostsim_lib.ptr(pts), C.byref(steps))
                assert (st, list(pts), steps.value) == tuple(o.leaf_rollout(SEED, g, step, 9, bool(det))), f"game {g} step {step} det {det}"
            m = o.allowed()
            legal = [a for a in range(39) if (m >> a) & 1]
            # weighted like rs-full-doko-cmd: mostly cards / no-announcement, sometimes calls
            a = int(prng.choice(legal))
            if kind == "no_solo" and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 and g % 2 == 0 else 24
            o.play(a)
            step += 1
    assert n_states > 5000
    assert n_dead == 0, f"{n_dead} dead ends"
