"""include/doko_env.hpp — the C++ host-side mirror of the reference's env traits (McEnvState / AzEnvState) — compiled (tests/cpp/env_check,
built by __graft_entry__.build()) and run on the GPU: a batch (new_game, random_rollout, encode_into_memory, allowed actions, summary)
and one game driven BY VALUE through the full trait method set, everything compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Fdo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xD0C05EED


def test_cpp_env_batch_and_by_value_state(orc, tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "env_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    n = 4096
    out = tmp_path / "env.bin"
    r = subprocess.run([exe, str(out), str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    off = [0]

    def take(dtype, count):
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off[0])
        off[0] += a.nbytes
        return a

    assert int(take("<u8", 2)[0]) == n
    pts, steps, obs, mask = take("<i4", n * 4).reshape(n, 4), take("<u4", n), take("<i8", n * 311).reshape(n, 311), take("<u8", n)
    recs = take(DK_STATE_DTYPE, n)
    stats = take("<u8", 270)
    # batch: fresh games 1000.. dealt with epoch 3, rolled out with announcements on epoch 4
    for i in range(0, n, 61):
        o = Fdo.new_game_philox(orc, SEED, 1000 + i, 3)
        assert o.export().tobytes() == recs[i].tobytes()
        assert np.array_equal(o.encode_pi(), obs[i]) and o.allowed() == int(mask[i])
        p, s = o.rollout(SEED, 1000 + i, 0, 4, with_announcements=True)
        assert list(pts[i]) == p and int(steps[i]) == s
    assert int(stats[0]) == n and int(stats[1]) == int(steps.sum()) and [int(x) for x in stats[2:6].view("<i8")] == [int(x) for x in pts.sum(0)]
    # one game by value
    n_walk, n_after, n_tok = (int(x) for x in take("<u8", 3))
    walk, after, tokens = take("<i8", n_walk * 8).reshape(n_walk, 8), take(DK_STATE_DTYPE, n_after), take("<i8", n_tok * 311).reshape(n_tok, 311)
    rew, rew8, roll = take("<f8", 4), take("<f4", 4), take("<f8", 4)
    o = Fdo.new_game_philox(orc, SEED, 1000, 3)
    calls = 0x1F << 33
    last = 0xFF
    tok_i = 0
    for k in range(n_walk):
        a, cur, n_young, n_old, id_lo, id_hi, m_first, m_below = (int(x) for x in walk[k])
        legal = o.allowed()
        assert cur == o.info()["current_player"]
        assert n_young == bin(legal & ~calls).count("1") and n_old == bin(legal).count("1")
        assert (id_hi << 32 | id_lo) == oracle_lib.fx_hash_record(o.export().tobytes(), last)
        assert m_first & ((1 << 64) - 1) == int(orc.orc_fdo_mc_allowed(o.h, 1)) and m_below & ((1 << 64) - 1) == int(orc.orc_fdo_mc_allowed(o.h, 0))
        o.play(a)
        last = a
        assert o.export().tobytes() == after[k].tobytes(), k
        if k % 9 == 0:
            assert np.array_equal(o.encode_pi(), tokens[tok_i])
            tok_i += 1
    info = o.info()
    assert info["phase"] == 3 and list(rew) == [float(x) for x in info["points"]] and list(rew8) == [np.float32(x) / np.float32(8) for x in info["points"]]
    o2 = Fdo.new_game_philox(orc, SEED, 1002, 3)
    p2, _ = o2.rollout(SEED, 77, 0, 5, with_announcements=False)
    assert list(roll) == [float(x) for x in p2]
