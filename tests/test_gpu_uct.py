"""GPU parity of dk_uct_search through the C ABI against oracle/mcts.hpp: visit counts, f32 values and chosen move of
every tree bit for bit, with and without determinization, and the complete mcts_cap_* decision (search per sample → PolicyFusion)."""
import numpy as np
import pytest

import oracle_lib
from test_uct_search import states_at_random_depth

pytestmark = pytest.mark.gpu
SEED = 0xAC75


@pytest.fixture(scope="module")
def dk():
    import master_doko_reinforcement_learning_b200 as pkg

    return pkg.DokoCuda(0)


def to_dev(objs):
    import torch

    from oracle_lib import DK_STATE_DTYPE
    recs = np.array([o.export() for o in objs], dtype=DK_STATE_DTYPE)
    return torch.from_numpy(np.frombuffer(recs.tobytes(), dtype=np.uint8).reshape(len(objs), 128).copy()).cuda()


@pytest.mark.parametrize("n,T,iterations,c,det,first_sub", [(70, 1, 300, 1.4, False, 0), (40, 3, 150, 0.7, True, 0), (6, 2, 2000, 2.0, True, 5), (33, 4, 1, 1.4, True, 0)])
def test_search_matches_oracle(dk, orc, n, T, iterations, c, det, first_sub):
    import torch

    objs = states_at_random_depth(orc, n, 31 * n + iterations)
    first_id = 12_000
    visits, values, action, status = dk.uct_search(to_dev(objs), iterations, c, dk.rng(SEED, first_id, 3, first_sub), trees_per_root=T, determinize=det)
    torch.cuda.synchronize()
    visits, values, action, status = visits.cpu().numpy().view(np.uint32), values.cpu().numpy(), action.cpu().numpy(), status.cpu().numpy()
    for i, o in enumerate(objs):
        for d in range(T):
            st_o, vis_o, val_o, act_o = o.uct_search(SEED, first_id + i, first_sub + d, iterations, c, 3, det)
            assert int(status[i, d]) == st_o, (i, d)
            assert (visits[i, d] == vis_o).all(), (i, d, visits[i, d], vis_o)
            assert (values[i, d].view(np.uint32) == val_o.view(np.uint32)).all(), (i, d)
            assert int(action[i, d]) == (act_o if act_o >= 0 else 0xFF)


def test_mcts_cap_decision_end_to_end(dk, orc):
    """DefaultImpiPolicy with CAPSampling + MCTS per sample + PolicyFusion (all_policies.rs:132-170): device == oracle."""
    import torch

    n, samples, iterations, c = 48, 8, 120, 1.4
    objs = [o for o in states_at_random_depth(orc, n + 16, 99, max_depth=60) if o.allowed()][:n]
    dev = to_dev(objs)
    visits, _, _, status = dk.uct_search(dev, iterations, c, dk.rng(SEED, 0, 9), trees_per_root=samples, determinize=True)
    allowed = dk.legal_mask(1, dev)
    decisions = {s: dk.fuse(s, visits, allowed, status)[0].cpu().numpy() for s in (0, 1)}
    torch.cuda.synchronize()
    for i, o in enumerate(objs):
        rows = []
        for d in range(samples):
            st_o, vis_o, _, _ = o.uct_search(SEED, i, d, iterations, c, 9, True)
            if st_o == 0:
                rows.append(vis_o)
        for s in (0, 1):
            assert int(decisions[s][i]) == oracle_lib.fuse(orc, s, np.array(rows), o.allowed()), (i, s)
