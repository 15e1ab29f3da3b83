"""PIMC move decision on the CPU: the oracle's PolicyFusion restatement against the reference's own four tests
(rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:125-313 → tests/golden/policy_fusion_cases.json), the device fuse logic
(hostsim) against the oracle on random rows with ties / zeros / failed samples, the additive root statistics against the fuse,
and the flat Monte-Carlo evaluator's per-thread logic against oracle/pimc.hpp."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import hostsim_lib
import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "policy_fusion_cases.json")))["cases"]
STRATEGY = {"PolicyFusionMaxN": 0, "PolicyFusionAverageStrategy": 1}
SEED = 0xF05E


def mask_of(actions):
    m = 0
    for a in actions:
        m |= 1 << a
    return m


def sim_fuse(sim, strategy, visits, allowed, status=None):
    v = np.ascontiguousarray(visits, dtype=np.uint32).reshape(-1, 39)
    n_ok = C.c_uint32()
    st = None if status is None else np.ascontiguousarray(status, dtype=np.uint8)
    a = sim.sim_fuse(strategy, hostsim_lib.ptr(v), hostsim_lib.ptr(st) if st is not None else None, v.shape[0], allowed, C.byref(n_ok))
    return a, n_ok.value


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_fusion_tests(orc, case):
    """The reference's known answers: oracle restatement and device logic."""
    strategy, allowed = STRATEGY[case["strategy"]], mask_of(case["allowed"])
    assert oracle_lib.fuse(orc, strategy, case["visits"], allowed) == case["expected"]
    sim = hostsim_lib.load()
    assert sim_fuse(sim, strategy, case["visits"], allowed) == (case["expected"], len(case["visits"]))


def random_rows(prng, n_rows, allowed, kind):
    legal = [a for a in range(39) if (allowed >> a) & 1]
    v = np.zeros((n_rows, 39), dtype=np.uint32)
    for r in range(n_rows):
        if kind == "ties":
            v[r, legal] = prng.integers(0, 3, size=len(legal))
        elif kind == "flat_mc":                       # rows summing to a power of two, like n_rollouts = 32
            picks = prng.choice(legal, size=32)
            for a in picks:
                v[r, a] += 1
        elif kind == "wide":
            v[r, legal] = prng.integers(0, 1 << 31, size=len(legal))
        else:                                         # "stray": visits on actions that are not allowed, some all-zero rows (0/0 = NaN)
            v[r] = prng.integers(0, 50, size=39)
            if r % 5 == 4:
                v[r] = 0
    return v


@pytest.mark.parametrize("kind", ["ties", "flat_mc", "wide", "stray"])
def test_device_fuse_matches_oracle(orc, kind):
    sim = hostsim_lib.load()
    prng = np.random.default_rng(5)
    for trial in range(300):
        n_legal = int(prng.integers(1, 13))
        allowed = mask_of(prng.choice(39, size=n_legal, replace=False).tolist())
        n_rows = int(prng.integers(1, 40))
        v = random_rows(prng, n_rows, allowed, kind)
        status = (prng.random(n_rows) < 0.2).astype(np.uint8)
        ok = v[status == 0]
        for strategy in (0, 1):
            got, n_ok = sim_fuse(sim, strategy, v, allowed, status)
            assert n_ok == len(ok)
            if len(ok):
                assert got == oracle_lib.fuse(orc, strategy, ok, allowed), (kind, trial, strategy)
            got_all, n_all = sim_fuse(sim, strategy, v, allowed)
            assert n_all == n_rows and got_all == oracle_lib.fuse(orc, strategy, v, allowed), (kind, trial, strategy)


def test_root_stats_are_additive_and_decide_like_the_fuse(orc):
    """Split the rows over 'ranks', accumulate the statistics per part, sum them: MaxN decision == fuse over all rows; Average ==
    fuse_average when every row total is the same power of two (exact f32)."""
    sim = hostsim_lib.load()
    prng = np.random.default_rng(6)
    for trial in range(200):
        n_legal = int(prng.integers(1, 13))
        allowed = mask_of(prng.choice(39, size=n_legal, replace=False).tolist())
        n_rows = int(prng.integers(2, 48))
        v = random_rows(prng, n_rows, allowed, "flat_mc")
        status = (prng.random(n_rows) < 0.15).astype(np.uint8)
        total = np.zeros(80, dtype=np.int64)
        cut = sorted(prng.integers(0, n_rows + 1, size=3).tolist())
        for lo, hi in zip([0] + cut, cut + [n_rows]):
            part = np.zeros(80, dtype=np.int64)
            if hi > lo:
                pv, ps = np.ascontiguousarray(v[lo:hi]), np.ascontiguousarray(status[lo:hi])
                sim.sim_root_stats(hostsim_lib.ptr(pv), hostsim_lib.ptr(ps), hi - lo, allowed, hostsim_lib.ptr(part))
            total += part
        ok = v[status == 0]
        assert total[78] == len(ok) and total[79] == 0
        assert (total[39:78] == ok.sum(axis=0)).all()
        for strategy in (0, 1):
            pick = sim.sim_root_pick(strategy, hostsim_lib.ptr(total), allowed)
            if len(ok) == 0:
                assert pick == 0xFF
            else:
                assert pick == oracle_lib.fuse(orc, strategy, ok, allowed), (trial, strategy)


def test_flat_mc_device_logic_matches_oracle(orc):
    """fdo_pimc_kernel's per-thread program, run sequentially on the CPU, against oracle/pimc.hpp on states of every phase."""
    sim = hostsim_lib.load()
    prng = np.random.default_rng(8)
    n_checked = 0
    phases = set()
    for g in range(24):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        depth = int(prng.integers(0, 70))
        for _ in range(depth):
            if not o.allowed():
                break
            m = o.allowed()
            legal = [a for a in range(39) if (m >> a) & 1]
            o.play(int(prng.choice(legal)))
        rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
        phases.add(o.info()["phase"])
        for det in (0, 3):
            st_o, vis_o, val_o = o.flat_mc(SEED, 100 + g, det, 4, 2)
            vis_d = np.zeros(39, dtype=np.uint32)
            val_d = np.zeros(39, dtype=np.int64)
            st_d = sim.sim_fdo_flat_mc(hostsim_lib.ptr(rec), SEED, 100 + g, det, 4, 2, hostsim_lib.ptr(vis_d), hostsim_lib.ptr(val_d))
            assert st_d == st_o, (g, det)
            assert (vis_d == vis_o).all() and (val_d == val_o).all(), (g, det, vis_d, vis_o, val_d, val_o)
            if st_o == 0 and o.allowed():
                assert vis_o.sum() == 4
                m = o.allowed()
                assert all(vis_o[a] == 0 and val_o[a] == 0 for a in range(39) if not (m >> a) & 1)
            n_checked += 1
    assert n_checked == 48 and len(phases) >= 3
