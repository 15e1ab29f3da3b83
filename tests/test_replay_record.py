"""The reference's on-disk experience record (bincode DBRecord, experience_replay_buffer3.rs:11-20,94-121) — oracle restatement
against a hand-assembled known answer of the published bincode 1.x format, and the GPU packer against the oracle."""
import struct

import numpy as np
import pytest

import oracle_lib


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
