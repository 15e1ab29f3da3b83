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
