#!/usr/bin/env python3
"""K5 (lock-step env step + 311-token rows) and the bare encoder at 2^22 games with i64, int32 and uint8 rows: ms, rows/s, GB/s against
the measured HBM peak.  Prints one JSON object (profiles/r02_narrow_rows.json)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
SEED = 0xD0C05EED


def main():
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    dk = pkg.DokoCuda(0)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6549.4)
    n = 1 << 22
    st0 = dk.new_games(pkg.DK_FDO, n, dk.rng(SEED, 0, 5))
    for k in range(24):
        dk.step_random_encode(st0, dk.rng(SEED, 0, k), want_obs=False)
    act = torch.empty((n,), dtype=torch.uint8, device="cuda")

    def timed(fn, iters=5, warm=3):
        for k in range(warm):
            fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            fn(warm + k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"games": n, "hbm_peak_gbs": hbm}
    for name, dt, nb in (("i64", torch.int64, 8), ("int32", torch.int32, 4), ("uint8", torch.uint8, 1)):
        st = st0.clone()
        obs = torch.empty((n, 311), dtype=dt, device="cuda")
        if dt == torch.int64:
            ms = timed(lambda k: dk.step_random_encode(st, dk.rng(SEED, 0, 100 + k), obs_out=obs, action_out=act))
            ms_e = timed(lambda k: dk.encode(pkg.DK_LAYOUT_FDO_PI311, st, out=obs))
        else:
            ms = timed(lambda k: dk.step_random_encode_narrow(st, dk.rng(SEED, 0, 100 + k), obs_out=obs, action_out=act))
            ms_e = timed(lambda k: dk.encode_narrow(pkg.DK_LAYOUT_FDO_PI311, st, out=obs))
        b5, be = 256 + 311 * nb, 128 + 311 * nb
        out[name] = {"step_encode_ms": ms, "step_encodes_per_s": n / ms * 1e3, "step_encode_hbm_frac": b5 * n / ms / 1e6 / hbm, "bytes_per_game": b5,
                     "encode_ms": ms_e, "encodes_per_s": n / ms_e * 1e3, "encode_hbm_frac": be * n / ms_e / 1e6 / hbm}
        del obs, st
    print(json.dumps(out))


if __name__ == "__main__":
    main()
