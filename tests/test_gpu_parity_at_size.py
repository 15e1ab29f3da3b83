"""Full-size parity: the CUDA path against the CPU oracle, bit for bit, on EVERY unit of large batches —
2^20 fresh games per engine / policy, 2^20 card_matching samples (BASELINE config-3 info-states and info-states at every stage of a
game, the shape of the reference's own soak rs-full-doko-cmd/src/main.rs:190-283: reservation-, announcement- and card-phase states,
every sample through is_consistent), 2^18 sample_assignment samples, all 8192 config-4 leaves x 64 rollouts, lock-step env steps with
the 311-token rows of 2^16 games per step.  The oracle runs on all host threads (oracle/capi.cpp orc_bulk_*, pinned to the
one-at-a-time oracle functions by tests/test_oracle_bulk.py)."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import Bulk

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stats_from(points, steps):
    """dk_playout_stats of a batch computed with numpy from per-game results."""
    p = points.astype(np.int64)
    hist = np.bincount(np.minimum(steps.astype(np.int64), 255), minlength=256)
    return {"games": len(steps), "game_steps": int(steps.astype(np.int64).sum()), "point_sum": [int(x) for x in p.sum(0)],
            "point_sq_sum": [int(x) for x in (p * p).sum(0)], "wins": [int(x) for x in (p > 0).sum(0)], "step_hist": hist.astype(np.uint64)}


@pytest.mark.parametrize("engine,ann", [(1, True), (1, False), (0, False)])
def test_fresh_playouts_2p20_every_game_and_every_result_form(dk, orc, engine, ann):
    """2^20 fresh games on a random (seed, first id, epoch): points and step counts of every game == oracle; the compact (5 B), packed
    (3 B) host forms decode to the same values; the device-reduced summary == the reduction of the oracle's per-game results."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 1 << 20
    prng = np.random.default_rng(1000 + 2 * engine + int(ann))
    seed, first_id, epoch = int(prng.integers(0, 1 << 62)), int(prng.integers(0, 1 << 40)), int(prng.integers(0, 1 << 20))
    flags = pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS if ann else 0
    rng = dk.rng(seed, first_id, epoch)
    pts, st = dk.playout(engine, n, rng, flags=flags)
    torch.cuda.synchronize()
    ref = oracle_lib.playout_philox(orc, engine, n, seed, first_id=first_id, epoch=epoch, with_announcements=ann)
    gp, gs = pts.cpu().numpy(), st.cpu().numpy().astype(np.uint32)
    assert np.array_equal(gp, ref["points"]) and np.array_equal(gs, ref["steps"])
    # host-transfer forms
    cp, cs = dk.playout_host_compact(engine, n, rng, flags=flags)
    assert np.array_equal(cp.astype(np.int32), ref["points"]) and np.array_equal(cs.astype(np.uint32), ref["steps"])
    pp, ps = dk.playout_host_packed(engine, n, rng, flags=flags)
    assert np.array_equal(pkg.api.unpack_points(pp), ref["points"]) and np.array_equal(ps.astype(np.uint32), ref["steps"])
    pp2, none = dk.playout_host_packed(engine, n, rng, flags=flags, want_steps=False)
    assert none is None and np.array_equal(pp2, pp)
    # summary, device and host entry points, and additivity over a split of the batch
    exp = stats_from(ref["points"], ref["steps"])
    got = dk.playout_summary_host(engine, n, rng, flags=flags).as_dict()
    for k in ("games", "game_steps", "point_sum", "point_sq_sum", "wins"):
        assert got[k] == exp[k], k
    assert np.array_equal(got["step_hist"], exp["step_hist"])
    cut = 333_333
    acc = dk.playout_summary(engine, cut, rng, flags=flags)
    dk.playout_summary(engine, n - cut, dk.rng(seed, first_id + cut, epoch), flags=flags, stats=acc, accumulate=True)
    torch.cuda.synchronize()
    w = acc.cpu().numpy()
    assert int(w[0]) == n and int(w[1]) == exp["game_steps"] and [int(x) for x in w[2:6]] == exp["point_sum"]
    assert np.array_equal(w[14:270].astype(np.uint64), exp["step_hist"])


def test_playout_summary_from_stored_states(dk, orc):
    """dk_playout_summary over stored mid-game records == reduction of dk_playout's per-game results (ragged size: partial last block)."""
    import torch

    b = Bulk(orc, 1, 10_007, SEED, first_id=0, epoch=3, mode=1)
    states = dev(b.bytes())
    rng = dk.rng(SEED, 5, 11)
    pts, st = dk.playout(1, b.n, rng, states=states, flags=1)
    acc = dk.playout_summary(1, b.n, rng, states=states, flags=1)
    torch.cuda.synchronize()
    exp = stats_from(pts.cpu().numpy(), st.cpu().numpy().astype(np.uint32))
    w = acc.cpu().numpy()
    assert int(w[0]) == b.n and int(w[1]) == exp["game_steps"] and [int(x) for x in w[2:6]] == exp["point_sum"]
    assert [int(x) for x in w[6:10]] == exp["point_sq_sum"] and [int(x) for x in w[10:14]] == exp["wins"]
    assert np.array_equal(w[14:270].astype(np.uint64), exp["step_hist"])


@pytest.mark.parametrize("mode,n,S", [(0, 4096, 256), (1, 16384, 64)])
def test_card_matching_2p20_samples_equal_oracle_and_are_consistent(dk, orc, mode, n, S):
    """mode 0: BASELINE config-3 info-states; mode 1: info-states at every stage of a game (reservation / announcement / card phase).
    Every sample == the oracle's card_matching on the same stream, none is a dead end, every one passes is_consistent."""
    import torch

    b = Bulk(orc, 1, n, SEED, first_id=1 << 33, epoch=3, mode=mode)
    phases = b.recs["meta"] & 3
    if mode == 1:
        assert (phases == 0).sum() > n // 50 and (phases == 1).sum() > n // 20 and (phases == 2).sum() > n // 20
    hands, res, status = dk.determinize(1, dev(b.bytes()), S, dk.rng(SEED, 1 << 33, 7, first_sub=5))
    torch.cuda.synchronize()
    h_o, r_o, s_o, cons, sec = b.determinize(S, epoch=7, first_sub=5)
    assert np.array_equal(status.cpu().numpy(), s_o)
    assert np.array_equal(hands.cpu().numpy().astype(np.uint64), h_o)
    assert np.array_equal(res.cpu().numpy(), r_o)
    assert int(s_o.max()) == 0 and int(np.abs(cons).max()) == 0, (int(s_o.max()), np.unique(cons))


def test_sample_assignment_2p18_samples_equal_oracle(dk, orc):
    import torch

    n, S = 4096, 64
    for mode in (0, 1):
        b = Bulk(orc, 0, n, SEED, first_id=77, epoch=0, mode=mode)
        hands, _, status = dk.determinize(0, dev(b.bytes()), S, dk.rng(SEED, 77, 6))
        torch.cuda.synchronize()
        h_o, _, s_o, _, sec = b.determinize(S, epoch=6)
        assert np.array_equal(status.cpu().numpy(), s_o) and np.array_equal(hands.cpu().numpy().astype(np.uint64), h_o), mode
        assert int(s_o.max()) == 0


@pytest.mark.parametrize("determinize", [True, False])
def test_all_8192_leaves_x_64_rollouts_equal_oracle(dk, orc, determinize):
    """BASELINE config 4's 8192 leaves: the exact integer point sums of every leaf == the oracle's, and the two halves of the rollouts
    (what two ranks would all-reduce) add up to the whole."""
    import torch

    n, R = 8192, 64
    b = Bulk(orc, 1, n, SEED, first_id=0, epoch=3, mode=0)
    states = dev(b.bytes())
    sums = dk.leaf_rollouts(states, R, dk.rng(SEED, 0, 9), determinize=determinize)
    a = dk.leaf_rollouts(states, R // 2, dk.rng(SEED, 0, 9, first_sub=0), determinize=determinize)
    c = dk.leaf_rollouts(states, R // 2, dk.rng(SEED, 0, 9, first_sub=R // 2), determinize=determinize)
    torch.cuda.synchronize()
    exp, sec = b.leaf_rollouts(R, epoch=9, determinize=determinize)
    assert np.array_equal(sums.cpu().numpy(), exp)
    assert torch.equal(a + c, sums)


@pytest.mark.parametrize("determinize", [True, False])
def test_leaves_at_every_game_stage_equal_oracle(dk, orc, determinize):
    """Leaves of the reference's soak recipe (reservation-, announcement- and card-phase states, most of them INSIDE a trick): the
    chained card draws of the running trick continue from the legal-card counts of the plays already made — for determinized rollouts
    counted on the SAMPLED hands (fdo_live_with_sample).  Exact integer point sums of every leaf == the oracle's."""
    import torch

    n, R = 2048, 48
    b = Bulk(orc, 1, n, SEED, first_id=1 << 20, epoch=3, mode=1)
    ci, phase = b.recs["card_index"], b.recs["meta"] & 3
    assert ((ci & 3) != 0).sum() > n // 3 and (phase == 0).sum() > 20 and (phase == 1).sum() > n // 20
    sums = dk.leaf_rollouts(dev(b.bytes()), R, dk.rng(SEED, 1 << 20, 9), determinize=determinize)
    torch.cuda.synchronize()
    exp, _ = b.leaf_rollouts(R, epoch=9, determinize=determinize)
    assert np.array_equal(sums.cpu().numpy(), exp)


def test_lockstep_steps_2p16_games_tokens_records_actions(dk, orc):
    """BASELINE config 5's check: after each lock-step env step the 311-token rows, the records and the actions of 2^16 games == oracle
    (games at every stage, so reservation, announcement, card and final steps are all in one batch); with and without skip_single."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    n = 1 << 16
    for skip in (False, True):
        b = Bulk(orc, 1, n, SEED, first_id=123, epoch=3, mode=1)
        states = dev(b.bytes())
        for k in range(12):
            flags = pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS | (0x100 if skip else 0)
            obs, act = dk.step_random_encode(states, dk.rng(SEED, 123, 200 + k), flags=flags)
            torch.cuda.synchronize()
            a_o, recs_o, obs_o = b.step(200 + k, True, skip, want_recs=True, want_obs=True)
            assert np.array_equal(act.cpu().numpy(), a_o), k
            assert states.cpu().numpy().tobytes() == recs_o.tobytes(), k
            assert np.array_equal(obs.cpu().numpy(), obs_o), k


def test_narrow_observation_rows_equal_the_i64_rows(dk, orc):
    """dk_encode_narrow / dk_step_random_encode_narrow write the SAME token values as the i64 entry points (which the tests above pin to
    the oracle), as int32 or uint8: every layout, games at every stage, ragged batch sizes (partial last block, image tails that are
    not a multiple of the 16-byte copy), and the env step leaves the same records and actions."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    for engine, layouts in ((1, (pkg.DK_LAYOUT_FDO_PI311,)), (0, (pkg.DK_LAYOUT_DO110, pkg.DK_LAYOUT_DO114))):
        for n in (1, 127, 129, 10_007, 1 << 16):
            b = Bulk(orc, engine, n, SEED, first_id=31, epoch=3, mode=1)
            states = dev(b.bytes())
            for layout in layouts:
                ref = dk.encode(layout, states)
                assert int(ref.max()) < 256 and int(ref.min()) >= 0
                for dt in (torch.int32, torch.uint8):
                    out = dk.encode_narrow(layout, states, dtype=dt)
                    assert out.dtype == dt and torch.equal(out.to(torch.int64), ref), (engine, n, layout, dt)
    n = 50_001
    b = Bulk(orc, 1, n, SEED, first_id=77, epoch=3, mode=1)
    for dt in (torch.int32, torch.uint8):
        st_a, st_b = dev(b.bytes()), dev(b.bytes())
        for k in range(6):
            flags = pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS | (0x100 if k & 1 else 0)
            obs, act = dk.step_random_encode(st_a, dk.rng(SEED, 77, 300 + k), flags=flags)
            obs_n, act_n = dk.step_random_encode_narrow(st_b, dk.rng(SEED, 77, 300 + k), flags=flags, dtype=dt)
            assert torch.equal(act, act_n) and torch.equal(st_a, st_b) and torch.equal(obs_n.to(torch.int64), obs), (dt, k)
    # refused: a misaligned output, an element size that is not 1 or 4
    buf = torch.empty(311 * 8 + 64, dtype=torch.uint8, device="cuda")
    st = dev(b.bytes())[:8].contiguous()
    rc = dk.L.dk_encode_narrow(dk.ctx, pkg.DK_LAYOUT_FDO_PI311, 1, 8, st.data_ptr(), buf.data_ptr() + 8, None)
    assert rc == 1                                                   # DK_ERR_INVALID_ARGUMENT
    rc = dk.L.dk_encode_narrow(dk.ctx, pkg.DK_LAYOUT_FDO_PI311, 2, 8, st.data_ptr(), buf.data_ptr(), None)
    assert rc == 1
    assert torch.equal(dk.encode_narrow(pkg.DK_LAYOUT_FDO_PI311, st).to(torch.int64), dk.encode(pkg.DK_LAYOUT_FDO_PI311, st))   # the context is still healthy


def test_az_mask_id_and_random_action_on_every_stage(dk, orc):
    """dk_legal_mask_az == FdoAzEnvState::allowed_actions_by_action_index / number_of_allowed_actions for every (is_secondary, epoch);
    dk_state_id == FxHasher64 over the record; dk_random_action == the action the lock-step step would play."""
    import torch

    n = 20_000
    b = Bulk(orc, 1, n, SEED, first_id=9, epoch=3, mode=1)
    states = dev(b.bytes())
    legal = dk.legal_mask(1, states).cpu().numpy().astype(np.uint64)
    calls = np.uint64(0x1F << 33)
    assert ((legal & calls) != 0).sum() > 100                       # states in which a call is allowed are in the batch
    for sec, epoch in ((False, 0), (False, 9), (False, 10), (True, 10), (True, 3), (False, 1 << 40)):
        mask, cnt = dk.legal_mask_az(states, sec, epoch)
        drop_m = sec or epoch < 10
        exp_m = legal & ~calls if drop_m else legal
        exp_c = legal & ~calls if epoch < 10 else legal
        assert np.array_equal(mask.cpu().numpy().astype(np.uint64), exp_m)
        assert np.array_equal(cnt.cpu().numpy(), np.array([bin(int(x)).count("1") for x in exp_c], dtype=np.uint8))
    for i in range(0, n, 997):                                       # against the oracle's restatement of the trait methods
        o = oracle_lib.Fdo.from_dk_state(orc, b.recs[i:i + 1])
        for sec, epoch in ((False, 0), (False, 10), (True, 10)):
            m = int(dk.legal_mask_az(states[i:i + 1], sec, epoch, want_count=False)[0][0]) & ((1 << 64) - 1)
            assert m == int(orc.orc_fdo_az_allowed(o.h, int(sec), epoch))
    last = torch.from_numpy((np.arange(n) % 40).astype(np.uint8)).cuda()
    ids = dk.state_id(states, last).cpu().numpy().astype(np.uint64)
    ids_none = dk.state_id(states).cpu().numpy().astype(np.uint64)
    raw = b.bytes()
    for i in list(range(0, n, 501)) + [n - 1]:
        assert int(ids[i]) == oracle_lib.fx_hash_record(raw[i], i % 40) and int(ids_none[i]) == oracle_lib.fx_hash_record(raw[i], 0xFF)
    assert len(np.unique(ids_none)) == len(np.unique(raw, axis=0))  # distinct records, distinct ids (no collision in 20 000)
    for ann in (True, False):
        act = dk.random_action(1, states, dk.rng(SEED, 9, 77), flags=1 if ann else 0).cpu().numpy()
        b2 = Bulk(orc, 1, n, SEED, first_id=9, epoch=3, mode=1)
        a_o, _, _ = b2.step(77, ann, False)
        assert np.array_equal(act, a_o)


def test_misaligned_and_strided_buffers_are_refused(dk):
    """a sliced view / offset buffer must give DK_ERR_INVALID_ARGUMENT instead of a sticky device fault."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    states = dk.new_games(1, 64, dk.rng(SEED, 0, 0))
    flat = torch.empty((64 * 128 + 8,), dtype=torch.uint8, device="cuda")
    off = flat[8:8 + 64 * 128].view(64, 128)                         # 8-byte offset: contiguous but not 16-byte aligned
    off.copy_(states)
    with pytest.raises(pkg.DokoCudaError):
        dk.legal_mask(1, off)
    with pytest.raises(pkg.DokoCudaError):
        dk.apply(1, off, torch.zeros(64, dtype=torch.uint8, device="cuda"))
    with pytest.raises(pkg.DokoCudaError):
        dk.legal_mask(1, states[::2])                                # strided view
    assert int(dk.legal_mask(1, states).min()) != 0                  # the context is still healthy
