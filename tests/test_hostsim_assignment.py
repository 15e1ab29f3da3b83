"""rs-doko-assignment: device logic (hostsim) vs the Vec-based oracle restatement, plus sample properties."""
import ctypes as C

import numpy as np

import hostsim_lib
from oracle_lib import DK_STATE_DTYPE, Doko

SEED = 4242


def popcount(x):
    return bin(x).count("1")


def test_sample_assignment_soak(orc):
    sim = hostsim_lib.load()
    prng = np.random.default_rng(3)
    n = dead = 0
    for g in range(60):
        o = Doko.new_game_philox(orc, SEED, g, 0)
        step = 0
        while o.allowed():
            rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
            real = o.hands()
            obs = o.info()["current_player"]
            for sample in range(3):
                st_o, h_o = o.sample_assignment(SEED, g * 100 + step, sample)
                hd = (C.c_uint64 * 4)()
                st_d = sim.sim_doko_assign(hostsim_lib.ptr(rec), SEED, g * 100 + step, sample, 0, hd)
                assert (st_d, [int(x) for x in hd]) == (st_o, h_o), f"game {g} step {step} sample {sample}"
                n += 1
                if st_o:
                    dead += 1
                    continue
                # properties: hand sizes kept, observer's hand kept, the card multiset is exactly the real one
                assert [popcount(x) for x in h_o] == [popcount(x) for x in real]
                any_ = lambda h: (h | (h >> 24)) & 0xFFFFFF
                both = lambda h: (h & (h >> 24)) & 0xFFFFFF
                assert any_(h_o[obs]) == any_(real[obs]) and both(h_o[obs]) == both(real[obs])
                cnt = lambda hs: [sum(((any_(h) >> c) & 1) + ((both(h) >> c) & 1) for h in hs) for c in range(24)]
                assert cnt(h_o) == cnt(real)
            m = o.allowed()
            legal = [a for a in range(26) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if g % 2 == 0 and (m >> 25) & 1:
                a = 25
            o.play(a)
            step += 1
    assert n > 9000
    print("dead ends:", dead, "of", n)
