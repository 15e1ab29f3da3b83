"""GPU parity: full-rules playouts through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu

SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


@pytest.mark.parametrize("with_ann", [True, False])
@pytest.mark.parametrize("first_id", [0, (1 << 32) - 1000, 123456789012])
def test_fdo_fresh_playout_matches_oracle(dk, orc, with_ann, first_id):
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 1 << 15
    flags = pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS if with_ann else 0
    pts, steps = dk.playout(pkg.DK_FDO, n, dk.rng(SEED, first_id, 2), flags=flags)
    torch.cuda.synchronize()
    ref = oracle_lib.playout_philox(orc, 1, n, SEED, first_id, 2, with_ann)
    assert np.array_equal(pts.cpu().numpy(), ref["points"])
    assert np.array_equal(steps.cpu().numpy().astype(np.uint32), ref["steps"])


def test_fdo_fresh_playout_ragged_and_empty(dk, orc):
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    for n in (0, 1, 31, 257, 1000):
        pts, steps = dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 7, 0), flags=1)
        torch.cuda.synchronize()
        ref = oracle_lib.playout_philox(orc, 1, n, SEED, 7, 0, True)
        assert np.array_equal(pts.cpu().numpy(), ref["points"].reshape(n, 4))
        assert np.array_equal(steps.cpu().numpy().astype(np.uint32), ref["steps"])


def test_fdo_host_entry_point_matches_device(dk):
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 50000
    rng = dk.rng(SEED, 99, 5)
    pts_d, steps_d = dk.playout(pkg.DK_FDO, n, rng, flags=1)
    torch.cuda.synchronize()
    pts_h, steps_h = dk.playout_host(pkg.DK_FDO, n, rng, flags=1)
    assert np.array_equal(pts_d.cpu().numpy(), pts_h)
    assert np.array_equal(steps_d.cpu().numpy().astype(np.uint32), steps_h)


def test_fdo_full_size_properties(dk):
    """BASELINE config 2 size (2^24 games): zero-sum points, plausible step counts, sharding invariance."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 1 << 24
    pts, steps = dk.playout(pkg.DK_FDO, n, dk.rng(SEED, 0, 2), flags=1)
    torch.cuda.synchronize()
    assert int(pts.sum(dim=1).abs().max()) == 0          # Doppelkopf scoring is zero-sum (solo: 3x vs 3 * -x)
    assert int(steps.min()) >= 52 and int(steps.max()) <= 52 + 44 * 48
    half = n // 2
    pts2, steps2 = dk.playout(pkg.DK_FDO, half, dk.rng(SEED, half, 2), flags=1)
    torch.cuda.synchronize()
    assert torch.equal(pts[half:], pts2) and torch.equal(steps[half:], steps2)


def test_fdo_host_compact_and_chunked_paths(dk, orc):
    """dk_playout_host / _compact cut the batch into 2^21-game chunks: results must equal the single-launch device path."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = (1 << 22) + 12345
    rng = dk.rng(SEED, 1 << 33, 8)
    pts_d, steps_d = dk.playout(pkg.DK_FDO, n, rng, flags=1)
    torch.cuda.synchronize()
    pts_h, steps_h = dk.playout_host(pkg.DK_FDO, n, rng, flags=1)
    pts_c, steps_c = dk.playout_host_compact(pkg.DK_FDO, n, rng, flags=1)
    pd, sd = pts_d.cpu().numpy(), steps_d.cpu().numpy()
    assert np.array_equal(pd, pts_h) and np.array_equal(sd.astype(np.uint32), steps_h)
    assert np.array_equal(pd, pts_c.astype(np.int32)) and np.array_equal(sd, steps_c.astype(np.int32))
    assert int(np.abs(pd).max()) < 128 and int(sd.max()) < 256


def test_host_playout_from_host_states(dk, orc):
    import master_doko_reinforcement_learning_b200 as pkg
    from oracle_lib import DK_STATE_DTYPE, Fdo

    objs = [Fdo.new_game_philox(orc, SEED, i, 0) for i in range(64)]
    for i, o in enumerate(objs):
        for k in range(3 * i):
            o.random_step(SEED, i, 0, True, k)
    recs = np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE)
    pts, steps = dk.playout_host(pkg.DK_FDO, len(objs), dk.rng(SEED, 40, 2), states=recs, flags=0)
    for i, o in enumerate(objs):
        rp, rs = o.rollout(SEED, 40 + i, 0, 2, False)
        assert list(pts[i]) == rp and int(steps[i]) == rs
