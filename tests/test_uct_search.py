"""UCT search on the CPU: the device per-thread program (hostsim) against the oracle's restatement of the reference's
MCTS (oracle/mcts.hpp ← rs-doko-mcts/src/mcts/{node,mcts}.rs) — visit counts, f32 values and the chosen move bit for bit — plus
structural properties of the reference's search."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

SEED = 0xAC75


def sim_search(sim, rec, unit, sub, iterations, c, epoch=0, determinize=False):
    visits = np.zeros(39, dtype=np.uint32)
    values = np.zeros(39, dtype=np.float32)
    action = C.c_int32()
    st = sim.sim_fdo_uct_search(hostsim_lib.ptr(rec), SEED, unit, sub, epoch, int(determinize), iterations, c, hostsim_lib.ptr(visits),
                                hostsim_lib.ptr(values), C.byref(action))
    return st, visits, values, action.value


def states_at_random_depth(orc, n, seed, max_depth=75):
    prng = np.random.default_rng(seed)
    out = []
    for g in range(n):
        o = Fdo.new_game_philox(orc, SEED, 1000 + g, 0)
        depth = int(prng.integers(0, 4)) if g % 5 == 0 else int(prng.integers(0, max_depth))      # every 5th state sits in the reservation phase
        for _ in range(depth):
            m = o.allowed()
            if not m:
                break
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if g % 2 == 0 and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 else 24
            o.play(a)
        out.append(o)
    return out


def test_mc_allowed_actions_filter(orc):
    """McFullDokoEnvState::allowed_actions (env_state_full_doko.rs:132-172): device == oracle at every state of random games."""
    sim = hostsim_lib.load()
    prng = np.random.default_rng(4)
    n = 0
    for g in range(30):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        while o.allowed():
            rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
            for first in (0, 1):
                assert sim.sim_uct_allowed(hostsim_lib.ptr(rec), first) == orc.orc_fdo_mc_allowed(o.h, first)
            m = o.allowed()
            o.play(int(prng.choice([a for a in range(39) if (m >> a) & 1])))
            n += 1
    assert n > 1500


@pytest.mark.parametrize("iterations,c", [(1, 1.4), (40, 1.4), (300, 0.5), (300, 3.0), (1200, 1.4)])
def test_device_search_matches_oracle(orc, iterations, c):
    sim = hostsim_lib.load()
    objs = states_at_random_depth(orc, 14 if iterations > 1000 else 30, iterations)
    phases = set()
    for i, o in enumerate(objs):
        rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
        phases.add(o.info()["phase"])
        for det in (False, True):
            st_o, vis_o, val_o, act_o = o.uct_search(SEED, 50 + i, 2, iterations, c, 7, det)
            st_d, vis_d, val_d, act_d = sim_search(sim, rec, 50 + i, 2, iterations, c, 7, det)
            assert st_d == st_o, (i, det)
            assert (vis_d == vis_o).all(), (i, det, vis_d, vis_o)
            assert (val_d.view(np.uint32) == val_o.view(np.uint32)).all(), (i, det, val_d, val_o)
            assert act_d == act_o
            if st_o == 0 and o.allowed():
                m = o.allowed()
                n_legal = bin(m).count("1")
                # every iteration adds exactly one visit to one root child; children appear one per iteration until all are expanded
                assert vis_o.sum() == iterations
                assert (vis_o > 0).sum() == min(n_legal, iterations)
                assert all(vis_o[a] == 0 for a in range(39) if not (m >> a) & 1)
                assert vis_o[act_o] == vis_o.max()
    assert len(phases) >= 3


def test_search_prefers_the_winning_card(orc):
    """Sanity of the restated search: in the last trick with two cards left, perfect information, the move with the better outcome
    for the mover gets the larger share of the visits."""
    found = 0
    for g in range(200):
        o = Fdo.new_game_philox(orc, SEED, 5000 + g, 0)
        while o.allowed() and o.info()["card_index"] < 40:
            m = o.allowed()
            a = [x for x in range(39) if (m >> x) & 1 and x not in (33, 34, 35, 36, 37)][0]
            o.play(a)
        m = o.allowed()
        legal = [x for x in range(24) if (m >> x) & 1]
        if o.info()["phase"] != 2 or len(legal) != 2:
            continue
        mover = o.info()["current_player"]
        st, vis, val, act = o.uct_search(SEED, g, 0, 400, 1.4)
        # flat estimate of both moves with many rollouts
        est = {}
        for a in legal:
            c = o.clone(); c.play(a)
            est[a] = np.mean([c.rollout(SEED, g, r)[0][mover] for r in range(300)]) if c.allowed() else c.info()
        if abs(est[legal[0]] - est[legal[1]]) < 1.5:
            continue
        better = max(legal, key=lambda a: est[a])
        assert act == better, (g, est, vis[legal])
        found += 1
        if found >= 5:
            break
    assert found >= 3


def test_two_stage_selection_equals_all_f64_selection():
    """uct_find_best_child takes the decision with an f32 filter + exact evaluation of the survivors; it must pick the same child as
    evaluating every child in f64 (what the reference does), in particular on exact ties and on near ties below the f32 resolution.  (The all-f64 side also skips the
    two integer shortcuts — identical statistics, equal rational Q — so they are checked against the plain evaluation too.)"""
    sim = hostsim_lib.load()
    prng = np.random.default_rng(12)
    n_cases = 0
    for trial in range(40000):
        nch = int(prng.integers(1, 13))
        kind = trial % 9
        if kind == 0:                                    # random statistics
            vis = prng.integers(1, 2000, size=nch)
            win = np.array([int(prng.integers(-60, 61)) * int(v) + int(prng.integers(-v, v + 1)) for v in vis])
        elif kind == 1:                                  # exact ties: equal fractions with different denominators
            base_v = int(prng.integers(1, 50)); base_w = int(prng.integers(-30, 31)) * base_v
            mult = prng.integers(1, 4, size=nch)
            vis = base_v * mult; win = base_w * mult
            j = int(prng.integers(0, nch)); win[j] += int(prng.integers(-3, 4))
        elif kind == 2:                                  # near ties far below f32 resolution: huge visit counts, win differs by one
            v = int(prng.integers(1 << 18, 1 << 22))
            vis = np.full(nch, v); win = np.full(nch, int(prng.integers(-20, 21)) * v) + prng.integers(-2, 3, size=nch)
        elif kind == 3:                                  # tiny span
            v = int(prng.integers(1000, 100000))
            vis = np.full(nch, v) + prng.integers(0, 3, size=nch); win = vis * int(prng.integers(-10, 11)) + prng.integers(0, 2, size=nch)
        elif kind == 4:                                  # young nodes: one or two visits each, small integer results
            vis = prng.integers(1, 3, size=nch); win = np.array([int(prng.integers(-12, 13)) for _ in range(nch)]) * vis
        elif kind == 5:                                  # some unvisited children
            vis = prng.integers(0, 3, size=nch); win = prng.integers(-10, 11, size=nch) * vis
        elif kind == 6:                                  # Q equal, exploration decides: visit counts differ by one
            v = int(prng.integers(10, 5000)); q = int(prng.integers(-40, 41))
            vis = v + prng.integers(0, 2, size=nch); win = q * vis
        elif kind == 7:                                  # Q equal as rationals at the top of the visit range: adjacent huge visit counts
            v = int(prng.integers(1 << 20, (1 << 22) - 4)); q = int(prng.integers(-127, 128))
            vis = v + prng.integers(0, 3, size=nch); win = q * vis
        else:                                            # extremes of the value range
            vis = prng.integers(1, 1 << 20, size=nch); win = vis * prng.choice([-127, 127, 0, 126], size=nch)
        vis = np.ascontiguousarray(vis, dtype=np.uint32); win = np.ascontiguousarray(win, dtype=np.int64)
        parent = int(vis.sum()) + 1
        if parent >= 1 << 23:
            parent = int(vis.max()) + 1
        c = float(prng.choice([0.0, 1.0e-4, 1.0e-3, 0.1, 0.5, 1.4, 2.0, 4.0, 25.0]))
        r = sim.sim_uct_select_check(nch, hostsim_lib.ptr(vis), hostsim_lib.ptr(win), parent, c)
        assert (r & 255) == (r >> 8), (trial, kind, nch, vis, win, parent, c, r & 255, r >> 8)
        n_cases += 1
    assert n_cases == 40000
