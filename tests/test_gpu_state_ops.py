"""GPU parity of the state-record API (new games, legal mask, apply, terminal, encode, lock-step step+encode, playouts from
states) against the oracle, batch-wise through the C ABI."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Doko, Fdo

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def recs(t):
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=DK_STATE_DTYPE)


def oracle_recs(objs):
    return np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE)


@pytest.mark.parametrize("engine", [0, 1])
def test_new_games_and_lockstep_walk(dk, orc, engine):
    """Walk n games to the end with oracle-chosen random legal actions; compare every record, mask, encoding each step."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 192
    rng = dk.rng(SEED, 5000, 3)
    states = dk.new_games(engine, n, rng)
    cls = Fdo if engine == 1 else Doko
    objs = [cls.new_game_philox(orc, SEED, 5000 + i, 3) for i in range(n)]
    prng = np.random.default_rng(7)
    layouts = [pkg.DK_LAYOUT_FDO_PI311] if engine == 1 else [pkg.DK_LAYOUT_DO110, pkg.DK_LAYOUT_DO114]
    for step in range(400):
        torch.cuda.synchronize()
        assert recs(states).tobytes() == oracle_recs(objs).tobytes(), f"step {step}"
        masks = dk.legal_mask(engine, states).cpu().numpy().astype(np.uint64)
        exp = np.array([o.allowed() for o in objs], dtype=np.uint64)
        assert np.array_equal(masks, exp)
        if step % 5 == 0:
            for lay in layouts:
                enc = dk.encode(lay, states).cpu().numpy()
                if engine == 1:
                    ref = np.stack([o.encode_pi() for o in objs])
                else:
                    ref = np.stack([o.encode(lay == pkg.DK_LAYOUT_DO114) for o in objs])
                assert np.array_equal(enc, ref)
        done, pts = dk.terminal(engine, states)
        done = done.cpu().numpy()
        assert np.array_equal(done != 0, exp == 0)
        if (exp == 0).all():
            assert np.array_equal(pts.cpu().numpy(), np.array([o.info()["points"] for o in objs], dtype=np.int32))
            break
        acts = np.zeros(n, dtype=np.uint8)
        for i, m in enumerate(exp):
            m = int(m)
            if m == 0:
                acts[i] = 0        # illegal on a finished game → err flag, state unchanged
                continue
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if i % 3 == 0 and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 else 24      # every third game: no solos (reach Normal / Wedding)
            acts[i] = a
            objs[i].play(a)
        err = dk.apply(engine, states, torch.from_numpy(acts).cuda()).cpu().numpy()
        assert np.array_equal(err != 0, exp == 0)
    else:
        raise AssertionError("games did not finish")


def test_step_random_encode_matches_oracle(dk, orc):
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 160
    states = dk.new_games(1, n, dk.rng(SEED, 0, 0))
    objs = [Fdo.new_game_philox(orc, SEED, i, 0) for i in range(n)]
    for step in range(130):
        for variant, flags in enumerate((pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS, 0, 0x100)):
            if step % 3 != variant:
                continue
            obs, act = dk.step_random_encode(states, dk.rng(SEED, 0, 100 + step), flags=flags)
            torch.cuda.synchronize()
            exp_act = np.array([o.step_site(SEED, i, 100 + step, flags & 1, bool(flags & 0x100)) for i, o in enumerate(objs)])
            assert np.array_equal(act.cpu().numpy().astype(np.int64), np.where(exp_act < 0, 255, exp_act))
            assert recs(states).tobytes() == oracle_recs(objs).tobytes()
            assert np.array_equal(obs.cpu().numpy(), np.stack([o.encode_pi() for o in objs]))
    assert all(o.allowed() == 0 for o in objs)


@pytest.mark.parametrize("engine", [0, 1])
def test_playout_from_states_matches_oracle(dk, orc, engine):
    import torch

    n = 256
    states = dk.new_games(engine, n, dk.rng(SEED, 77, 1))
    cls = Fdo if engine == 1 else Doko
    objs = [cls.new_game_philox(orc, SEED, 77 + i, 1) for i in range(n)]
    prng = np.random.default_rng(3)
    # advance game i by a game-specific number of random steps (mid-reservation, mid-round, mid-trick, ...)
    for i, o in enumerate(objs):
        for _ in range(int(prng.integers(0, 70))):
            m = o.allowed()
            if m == 0:
                break
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if i % 2 == 0 and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 else 24
            o.play(a)
    st = torch.from_numpy(np.frombuffer(oracle_recs(objs).tobytes(), dtype=np.uint8).reshape(n, 128).copy()).cuda()
    for with_ann in ((True, False) if engine == 1 else (False,)):
        pts, steps = dk.playout(engine, n, dk.rng(SEED, 900, 4), states=st, flags=1 if with_ann else 0)
        torch.cuda.synchronize()
        pts, steps = pts.cpu().numpy(), steps.cpu().numpy()
        for i, o in enumerate(objs):
            if engine == 1:
                rp, rs = o.rollout(SEED, 900 + i, 0, 4, with_ann)
            else:
                c = Doko(orc, orc.orc_doko_clone(o.h))
                while c.random_step(SEED, 900 + i, 4) >= 0:
                    pass
                rp, rs = c.info()["points"], c.info()["n_play_actions"] - o.info()["n_play_actions"]
            assert list(pts[i]) == rp and int(steps[i]) == rs, f"game {i}"


def test_doko_trace_config1(dk, orc):
    """BASELINE config 1 outputs: 52 action ids, player_eyes, re_players, player_points per game — bit-exact."""
    import torch

    n = 1 << 16
    pts, trace, aux = dk.playout_trace(0, n, dk.rng(SEED, 0, 1))
    torch.cuda.synchronize()
    ref = oracle_lib.playout_philox(orc, 0, n, SEED, 0, 1, True, 0, want_aux=True, trace_stride=52)
    assert np.array_equal(pts.cpu().numpy(), ref["points"])
    assert np.array_equal(trace.cpu().numpy(), ref["trace"])
    aux = aux.cpu().numpy().astype(np.uint32)
    eyes = np.stack([(aux[:, 2] >> (8 * s)) & 255 for s in range(4)], 1)
    assert np.array_equal(eyes, ref["aux"][:, 4:8].astype(np.uint32))
    assert np.array_equal(aux[:, 1], ref["aux"][:, 1].astype(np.uint32))


def test_doko_one_million_playouts_config1(dk, orc):
    import torch

    n = 1_000_000
    pts, steps = dk.playout(0, n, dk.rng(SEED, 0, 1))
    torch.cuda.synchronize()
    ref = oracle_lib.playout_philox(orc, 0, n, SEED, 0, 1, True, 0)
    assert np.array_equal(pts.cpu().numpy(), ref["points"])
    assert int(steps.min()) == 52 and int(steps.max()) == 52


def test_argument_errors_are_status_codes_not_crashes(dk):
    """Error behaviour of the boundary: bad arguments come back as DK_ERR_INVALID_ARGUMENT (the reference would panic), n == 0 is a no-op."""
    import ctypes as C

    import torch

    import master_doko_reinforcement_learning_b200.api as api

    L, ctx = dk.L, dk.ctx
    states = dk.new_games(1, 4, dk.rng(1, 0, 0))
    rng = dk.rng(1, 0, 0)
    out = torch.empty((4, 311), dtype=torch.int64, device="cuda")
    mask = torch.empty((4,), dtype=torch.int64, device="cuda")
    p = api._ptr
    INVALID = 1
    assert L.dk_legal_mask(ctx, 7, 4, p(states), p(mask), None) == INVALID                       # unknown engine
    assert L.dk_legal_mask(ctx, 1, 4, None, p(mask), None) == INVALID                            # NULL states
    assert L.dk_encode(ctx, 2, 4, p(states), p(out), 310, None) == INVALID                       # row_stride < row length
    assert L.dk_encode(ctx, 9, 4, p(states), p(out), 311, None) == INVALID                       # unknown layout
    assert L.dk_playout(ctx, 5, 0, 4, None, C.byref(rng), p(mask), None, None) == INVALID
    assert L.dk_pimc_evaluate(ctx, 4, 2, 0, p(states), C.byref(rng), p(out), None, None, None) == INVALID       # zero rollouts
    assert L.dk_fuse(ctx, 5, 4, 2, p(out), None, p(mask), p(mask), None, None) == INVALID        # unknown strategy
    assert L.dk_uct_search(ctx, 4, 1, 0, 10, C.c_float(1.4), p(states), C.byref(rng), p(out), 16, None, None, None, None, None) == INVALID   # workspace too small
    assert L.dk_encode(ctx, 2, 0, p(states), p(out), 311, None) == 0                             # empty batch
    assert L.dk_legal_mask(ctx, 1, 0, p(states), p(mask), None) == 0
    assert b"" is not None and L.dk_last_error(ctx) is not None
    # the context is still healthy
    m = dk.legal_mask(1, states)
    torch.cuda.synchronize()
    assert int((m != 0).sum()) == 4


@pytest.mark.parametrize("n", [1, 2, 131, 257])
def test_doko_encode_dense_pairs_ragged_strided_unaligned(dk, orc, n):
    """rs-doko rows are written in sector-aligned PAIRS when the output is dense and 32-byte aligned: odd counts (last pair half
    empty), padded strides and unaligned outputs (row-by-row path) must all give the oracle's rows."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    states = dk.new_games(0, n, dk.rng(SEED, 900, 1))
    objs = [Doko.new_game_philox(orc, SEED, 900 + i, 1) for i in range(n)]
    for k in range(4 + 13):                                   # reservations + 13 cards: lowest legal action
        acts = np.zeros(n, dtype=np.uint8)
        for i, o in enumerate(objs):
            m = int(o.allowed())
            acts[i] = (m & -m).bit_length() - 1
            o.play(int(acts[i]))
        assert int(dk.apply(0, states, torch.from_numpy(acts).cuda()).sum()) == 0
    for lay, with_res, length in ((pkg.DK_LAYOUT_DO110, False, 110), (pkg.DK_LAYOUT_DO114, True, 114)):
        ref = np.stack([o.encode(with_res) for o in objs])
        dense = dk.encode(lay, states)
        assert dense.shape == (n, length) and np.array_equal(dense.cpu().numpy(), ref)
        padded = torch.full((n, length + 7), -1, dtype=torch.int64, device="cuda")
        dk.encode(lay, states, out=padded, row_stride=length + 7)
        assert np.array_equal(padded[:, :length].cpu().numpy(), ref) and bool((padded[:, length:] == -1).all())
        flat = torch.full((n * length + 3,), -1, dtype=torch.int64, device="cuda")
        shifted = flat[1:1 + n * length].view(n, length)        # 8-byte aligned only
        dk.encode(lay, states, out=shifted)
        assert np.array_equal(shifted.cpu().numpy(), ref) and int(flat[0]) == -1 and bool((flat[1 + n * length:] == -1).all())


@pytest.mark.parametrize("engine", [0, 1])
def test_tma_tiles_equal_cooperative_copies(dk, engine):
    """dk_legal_mask / dk_apply move record tiles with TMA tensor copies; DOKO_CUDA_NO_TMA=1 selects the cooperative-copy kernels.
    Both must give identical masks, error flags and records at every step of 100 003 games (ragged last tile) played to the end
    with random legal actions, illegal actions (finished games, deliberately wrong indices) included."""
    import os

    import torch

    n = 100_003
    a_states = dk.new_games(engine, n, dk.rng(SEED, 77, 9))
    b_states = a_states.clone()
    gen = torch.Generator(device="cuda").manual_seed(5)
    try:
        for step in range(400):
            os.environ.pop("DOKO_CUDA_NO_TMA", None)
            m_a = dk.legal_mask(engine, a_states)
            os.environ["DOKO_CUDA_NO_TMA"] = "1"
            m_b = dk.legal_mask(engine, b_states)
            assert torch.equal(m_a, m_b), f"masks differ at step {step}"
            if int((m_a != 0).sum()) == 0:
                break
            # a random legal action per game: the r-th set bit of the mask; every 17th game (rotating) gets action 38 - that (mostly illegal)
            bits = ((m_a.unsqueeze(1) >> torch.arange(39, device="cuda")) & 1).to(torch.int32)            # [n,39]
            cnt = bits.sum(1).clamp(min=1)
            r = (torch.rand(n, device="cuda", generator=gen) * cnt).to(torch.int32).clamp(max=cnt - 1)
            act = ((bits.cumsum(1) > r.unsqueeze(1)) & (bits != 0)).to(torch.int32).argmax(1).to(torch.uint8)
            act[step % 17::17] = 38 - act[step % 17::17]
            os.environ.pop("DOKO_CUDA_NO_TMA", None)
            e_a = dk.apply(engine, a_states, act)
            os.environ["DOKO_CUDA_NO_TMA"] = "1"
            e_b = dk.apply(engine, b_states, act)
            assert torch.equal(e_a, e_b), f"error flags differ at step {step}"
            assert torch.equal(a_states, b_states), f"records differ at step {step}"
        else:
            raise AssertionError("games did not finish")
    finally:
        os.environ.pop("DOKO_CUDA_NO_TMA", None)
    done, _ = dk.terminal(engine, a_states)
    assert int(done.sum()) == n
