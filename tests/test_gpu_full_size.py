"""BASELINE.json full-size configurations on one B200, checked through size-independent properties (zero-sum scoring, card
conservation, sharding invariance, determinism) plus oracle spot checks on samples."""
import numpy as np
import pytest

from oracle_lib import DK_STATE_DTYPE, Fdo

pytestmark = pytest.mark.gpu
SEED = 0xD0C05EED


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def popcount64(x):
    """SWAR popcount on an int64 torch tensor (values < 2^48)."""
    x = x - ((x >> 1) & 0x5555555555555555)
    x = (x & 0x3333333333333333) + ((x >> 2) & 0x3333333333333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F0F0F0F0F
    return (x * 0x0101010101010101) >> 56 & 0x7F


def midgame_batch(dk, n, steps, epoch0=0):
    states = dk.new_games(1, n, dk.rng(SEED, 0, 3))
    for k in range(steps):
        dk.step_random_encode(states, dk.rng(SEED, 0, epoch0 + k), want_obs=False)
    return states


def test_config5_lockstep_selfplay_2p22(dk, orc):
    """2^22 games in lock-step: env step + 311-token encode until every game is over; spot-check tokens and records against the oracle."""
    import torch

    n = 1 << 22
    states = dk.new_games(1, n, dk.rng(SEED, 0, 3))
    obs = torch.empty((n, 311), dtype=torch.int64, device="cuda")
    sample = np.arange(0, n, n // 512)[:512]
    objs = [Fdo.new_game_philox(orc, SEED, int(i), 3) for i in sample]
    sidx = torch.from_numpy(sample).cuda()
    for step in range(200):
        dk.step_random_encode(states, dk.rng(SEED, 0, 50 + step), obs_out=obs)
        for i, o in zip(sample, objs):
            o.step_site(SEED, int(i), 50 + step, True, False)
        if step % 16 == 0 or step > 90:
            got = obs[sidx].cpu().numpy()
            assert np.array_equal(got, np.stack([o.encode_pi() for o in objs])), f"step {step}"
            done, pts = dk.terminal(1, states)
            if bool(done.all()):
                break
    else:
        raise AssertionError("not all games finished")
    assert int(pts.sum(dim=1).abs().max()) == 0                     # zero-sum
    recs = np.frombuffer(states[sidx].cpu().numpy().tobytes(), dtype=DK_STATE_DTYPE)
    assert recs.tobytes() == np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE).tobytes()
    assert int(obs[:, 310].min()) == 3 and int(obs[:, 310].max()) == 3   # phase token Finished everywhere


def test_config3_determinize_64k_x_4096(dk, orc):
    """65 536 mid-game info-states x 4096 samples (2.7e8 samples): no dead ends, hand sizes and the card multiset are conserved,
    the observer keeps its hand, splitting the samples over two calls (first_sub) changes nothing."""
    import torch

    n, S = 1 << 16, 4096
    states = midgame_batch(dk, n, 24)
    hands, res, status = dk.determinize(1, states, S, dk.rng(SEED, 0, 7))
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    real = states[:, :32].contiguous().view(torch.int64).view(n, 4)
    meta = states[:, 124:128].contiguous().view(torch.int32).view(n)
    obs_seat = ((meta >> 2) & 3).long()
    assert torch.equal(popcount64(hands), popcount64(real)[:, None, :].expand(n, S, 4))
    h_any = (hands | (hands >> 24)) & 0xFFFFFF
    r_any = (real | (real >> 24)) & 0xFFFFFF
    union = h_any[..., 0] | h_any[..., 1] | h_any[..., 2] | h_any[..., 3]
    assert torch.equal(union, (r_any[:, 0] | r_any[:, 1] | r_any[:, 2] | r_any[:, 3])[:, None].expand(n, S))
    own = torch.gather(hands, 2, obs_seat[:, None, None].expand(n, S, 1))[..., 0]
    assert torch.equal(own, torch.gather(real, 1, obs_seat[:, None])[:, 0][:, None].expand(n, S))
    del h_any, union, own
    # sharding over the sample axis
    sub = states[:2048]
    h_a, r_a, s_a = dk.determinize(1, sub, S // 2, dk.rng(SEED, 0, 7, first_sub=0))
    h_b, r_b, s_b = dk.determinize(1, sub, S // 2, dk.rng(SEED, 0, 7, first_sub=S // 2))
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([h_a, h_b], 1), hands[:2048]) and torch.equal(torch.cat([r_a, r_b], 1), res[:2048])
    # oracle spot check incl. is_consistent
    recs = np.frombuffer(states[:4].cpu().numpy().tobytes(), dtype=DK_STATE_DTYPE)
    hh, rr = hands[:4].cpu().numpy().astype(np.uint64), res[:4].cpu().numpy()
    for i in range(4):
        o = Fdo.from_dk_state(orc, recs[i:i + 1])
        for s in (0, 1, 4095):
            st, h_o, r_o = o.card_matching(SEED, i, s, 7)
            assert st == 0 and [int(x) for x in hh[i, s]] == h_o and list(rr[i, s]) == r_o
            assert o.is_consistent(h_o, r_o) == 0


def test_config4_leaf_rollouts_8192_x_1024(dk, orc):
    import torch

    n, R = 8192, 1024
    states = midgame_batch(dk, n, 20)
    sums = dk.leaf_rollouts(states, R, dk.rng(SEED, 0, 9), determinize=True)
    torch.cuda.synchronize()
    assert int(sums.sum(dim=1).abs().max()) == 0                    # every rollout is zero-sum, so are the sums
    assert int(sums.abs().max()) <= R * 127
    half_a = dk.leaf_rollouts(states, R // 2, dk.rng(SEED, 0, 9, first_sub=0))
    half_b = dk.leaf_rollouts(states, R // 2, dk.rng(SEED, 0, 9, first_sub=R // 2))
    torch.cuda.synchronize()
    assert torch.equal(half_a + half_b, sums)                       # what the 8-GPU all-reduce adds up
    rec = np.frombuffer(states[:1].cpu().numpy().tobytes(), dtype=DK_STATE_DTYPE)
    o = Fdo.from_dk_state(orc, rec)
    exp = np.zeros(4, dtype=np.int64)
    for r in range(R):
        st, pts, _ = o.leaf_rollout(SEED, 0, r, 9, True)
        exp += np.array(pts) if st == 0 else 0
    assert np.array_equal(sums[0].cpu().numpy(), exp)


def test_config3_sample_assignment_64k_x_4096(dk, orc):
    """BASELINE configs[2] for the simplified engine (rs-doko-assignment): 65 536 mid-game info-states x 4096 samples = 2.7e8
    sample_assignment draws.  Every sample keeps the hand sizes and the card multiset, the observer keeps its hand, no dead ends,
    splitting the sample axis changes nothing; 64 info-states x 64 samples spread over the batch are compared with the oracle."""
    import torch

    from oracle_lib import Bulk

    n, S = 1 << 16, 4096
    b = Bulk(orc, 0, n, SEED, first_id=0, epoch=0, mode=0)
    states = torch.from_numpy(b.bytes()).cuda()
    hands, _, status = dk.determinize(0, states, S, dk.rng(SEED, 0, 6))
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    real = states[:, :32].contiguous().view(torch.int64).view(n, 4)
    meta = states[:, 124:128].contiguous().view(torch.int32).view(n)
    obs_seat = ((meta >> 2) & 3).long()
    assert torch.equal(popcount64(hands), popcount64(real)[:, None, :].expand(n, S, 4))
    # the multiset: per card type the number of copies over the four hands (copy A bits 0-23, copy B bits 24-47)
    cnt = lambda h: ((h & 0xFFFFFF).unsqueeze(-1) >> torch.arange(24, device="cuda") & 1).sum(-2) + ((h >> 24 & 0xFFFFFF).unsqueeze(-1) >> torch.arange(24, device="cuda") & 1).sum(-2)
    sub = torch.arange(0, n, 257, device="cuda")
    assert torch.equal(cnt(hands[sub][:, ::64]), cnt(real[sub])[:, None, :].expand(len(sub), S // 64, 24))
    own = torch.gather(hands, 2, obs_seat[:, None, None].expand(n, S, 1))[..., 0]
    # (the sampler returns the observer's hand in canonical copy order, hand_from_vec(hand_to_vec(own)): copy A filled first)
    canon = lambda h: (((h | (h >> 24)) & 0xFFFFFF) | (((h & (h >> 24)) & 0xFFFFFF) << 24))
    assert torch.equal(own, canon(torch.gather(real, 1, obs_seat[:, None])[:, 0])[:, None].expand(n, S))
    del own
    h_a, _, _ = dk.determinize(0, states[:1024], S // 2, dk.rng(SEED, 0, 6, first_sub=0))
    h_b, _, _ = dk.determinize(0, states[:1024], S // 2, dk.rng(SEED, 0, 6, first_sub=S // 2))
    assert torch.equal(torch.cat([h_a, h_b], 1), hands[:1024])
    pick = np.arange(0, n, n // 64)[:64]
    small = Bulk(orc, 0, n, SEED, first_id=0, epoch=0, mode=0)         # (the recipe is deterministic: same states)
    h_o, _, s_o, _, _ = small.determinize(64, epoch=6)
    got = hands[torch.from_numpy(pick).cuda()][:, :64].cpu().numpy().astype(np.uint64)
    assert np.array_equal(got, h_o[pick]) and int(s_o[pick].max()) == 0
