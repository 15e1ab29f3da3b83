"""GPU parity of dk_encode_ipi through the C ABI: the reference's known answer and oracle rows on random games."""
import numpy as np
import pytest

from oracle_lib import DK_STATE_DTYPE, Fdo
from test_encode_ipi import ipi_case, partial_guess

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def dev(a):
    import torch

    return torch.from_numpy(a).cuda()


def test_reference_vector(dk):
    rec, hands, res, nxt, expected = ipi_case()
    states = dev(np.frombuffer(rec.tobytes(), dtype=np.uint8).reshape(1, 128).copy())
    out, err = dk.encode_ipi(states, dev(np.array([hands], dtype=np.uint64).view(np.int64)), dev(np.array([res], dtype=np.uint8)),
                             dev(np.array([nxt], dtype=np.uint8)))
    assert int(err[0]) == 0 and np.array_equal(out.cpu().numpy()[0], expected)


def test_random_games_match_oracle(dk, orc):
    prng = np.random.default_rng(9)
    recs, hands, ress, nxts, want = [], [], [], [], []
    for g in range(48):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        step = 0
        while o.allowed():
            if step % 3 == g % 3:
                a, r, nx = partial_guess(prng, o)
                recs.append(o.export()); hands.append(a); ress.append(r); nxts.append(nx)
                want.append(o.encode_ipi(a, r, nx))
            m = o.allowed()
            legal = [x for x in range(39) if (m >> x) & 1]
            act = int(prng.choice(legal))
            if g % 2 == 0 and (m >> 24) & 1:
                act = 25 if (m >> 25) & 1 else 24
            o.play(act)
            step += 1
    n = len(recs)
    assert n > 1000 and n % 128 != 0                                   # ragged last block
    rec = np.array(recs, dtype=DK_STATE_DTYPE)
    states = dev(np.frombuffer(rec.tobytes(), dtype=np.uint8).reshape(n, 128).copy())
    out, err = dk.encode_ipi(states, dev(np.array(hands, dtype=np.uint64).view(np.int64)), dev(np.array(ress, dtype=np.uint8)),
                             dev(np.array(nxts, dtype=np.uint8)), row_stride=320)
    assert int(err.max()) == 0
    got = out.cpu().numpy()[:, :311]
    assert np.array_equal(got, np.array(want))


def test_oversized_guess_is_flagged(dk, orc):
    o = Fdo.new_game_philox(orc, SEED, 5, 0)
    for _ in range(30):
        m = o.allowed()
        o.play([a for a in range(39) if (m >> a) & 1][0])
    rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
    cur = o.info()["current_player"]
    hands = np.zeros((1, 4), dtype=np.uint64)
    hands[0, (cur + 1) % 4] = (1 << 24) - 1                              # 24 guessed cards
    states = dev(np.frombuffer(rec.tobytes(), dtype=np.uint8).reshape(1, 128).copy())
    _, err = dk.encode_ipi(states, dev(hands.view(np.int64)), dev(np.full((1, 4), 0xFF, dtype=np.uint8)), dev(np.zeros(1, dtype=np.uint8)))
    assert int(err[0]) == 1
