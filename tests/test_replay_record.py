"""The reference's on-disk experience record (bincode DBRecord, experience_replay_buffer3.rs:11-20,94-121).  The reference holds no
golden bytes for it, so the record is pinned to an INDEPENDENT encoder of the bincode 1.x wire format written from the format's
specification (tests/golden/bincode_v1.py: generic struct / seq / scalar rules, nothing DBRecord-specific) and to the bytes that
encoder produced for three committed rows (tests/golden/dbrecord_golden.json, made by make_dbrecord_golden.py); the oracle restatement
and the GPU packer must both reproduce them."""
import base64
import json
import os
import struct
import sys

import numpy as np
import pytest

import oracle_lib


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_rows():
    rows = json.load(open(os.path.join(GOLDEN, "dbrecord_golden.json")))["rows"]
    st = np.array([r["state"] for r in rows], dtype=np.int64)
    va = np.array([r["value"] for r in rows], dtype=np.float32)
    po = np.array([r["policy"] for r in rows], dtype=np.float32)
    want = np.stack([np.frombuffer(base64.b64decode(r["bincode_b64"]), dtype=np.uint8) for r in rows])
    return st, va, po, want


def test_independent_bincode_encoder_on_known_answers():
    """The spec-driven encoder itself, on values whose bincode 1.x bytes are documented in the crate's README / tests: a (u32, String)
    style struct, a Vec<u16>, an Option — so that a mistake in the generic rules would not hide behind DBRecord's regular shape."""
    sys.path.insert(0, GOLDEN)
    import bincode_v1 as b

    assert b.encode(("struct", [("u32", 1), ("string", "ab")])) == bytes([1, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0x61, 0x62])
    assert b.encode(("seq", "u16", [1, 2, 3])) == bytes([3, 0, 0, 0, 0, 0, 0, 0, 1, 0, 2, 0, 3, 0])
    assert b.encode(("option", ("i64", -2))) == bytes([1]) + (-2).to_bytes(8, "little", signed=True)
    assert b.encode(("option", None)) == b"\x00"
    assert b.encode(("f32", 1.0)) == bytes([0, 0, 0x80, 0x3F])


def test_oracle_reproduces_the_golden_records(orc):
    st, va, po, want = golden_rows()
    sys.path.insert(0, GOLDEN)
    import bincode_v1

    for r in range(len(st)):                                                # the committed bytes are what the encoder gives today
        assert bincode_v1.db_record(st[r], va[r], po[r]) == want[r].tobytes()
    got = oracle_lib.replay_records(orc, st, va, po)
    assert np.array_equal(got, want)


def test_oracle_record_layout(orc):
    prng = np.random.default_rng(1)
    st = prng.integers(-5, 60, size=(3, 311)).astype(np.int64)
    va = prng.standard_normal((3, 4)).astype(np.float32)
    po = prng.random((3, 39)).astype(np.float32)
    got = oracle_lib.replay_records(orc, st, va, po)
    assert got.shape == (3, 2684)
    for r in range(3):
        want = struct.pack("<Q311q", 311, *st[r].tolist()) + struct.pack("<Q", 4) + va[r].tobytes() + struct.pack("<Q", 39) + po[r].tobytes()
        assert len(want) == 2684 and got[r].tobytes() == want


@pytest.mark.gpu
def test_gpu_packer_matches_oracle(orc):
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    dk = pkg.DokoCuda(0)
    n = 64
    states = dk.new_games(1, n, dk.rng(7, 0, 0))
    sp = dk.self_play(n, n * 130)
    for t in range(260):
        sp.begin_turn(states, 0, 1.0, dk.rng(7, 0, t))
        sp.uniform_search(dk.rng(7, 0, t))
        sp.end_turn(states)
    sp.finalize(states)
    rows, dropped, unfinished = sp.counts()
    assert rows > 60 * n and dropped == 0 and unfinished == 0
    gst, gva, gpo, gwant = golden_rows()                                        # the committed golden records through the GPU packer
    grec = dk.pack_replay_records(torch.from_numpy(gst).cuda(), torch.from_numpy(gva).cuda(), torch.from_numpy(gpo).cuda())
    assert np.array_equal(grec.cpu().numpy(), gwant)
    rec = dk.pack_replay_records(sp.states[:rows], sp.value[:rows], sp.policy[:rows])
    torch.cuda.synchronize()
    want = oracle_lib.replay_records(orc, sp.states[:rows].cpu().numpy(), sp.value[:rows].cpu().numpy(), sp.policy[:rows].cpu().numpy())
    assert np.array_equal(rec.cpu().numpy(), want)
    # a row count that is not a multiple of four (tail rows go word by word) and an output that is only 4-byte aligned (no 16-byte stores)
    m = rows - (rows % 4) - 1
    rec2 = dk.pack_replay_records(sp.states[:m], sp.value[:m], sp.policy[:m])
    assert np.array_equal(rec2.cpu().numpy(), want[:m])
    raw = torch.zeros((m * 2684 + 16,), dtype=torch.uint8, device="cuda")
    out3 = raw[4:4 + m * 2684].view(m, 2684)
    dk.pack_replay_records(sp.states[:m], sp.value[:m], sp.policy[:m], out=out3)
    torch.cuda.synchronize()
    assert np.array_equal(out3.cpu().numpy(), want[:m]) and int(raw[:4].sum()) == 0 and int(raw[4 + m * 2684:].sum()) == 0
