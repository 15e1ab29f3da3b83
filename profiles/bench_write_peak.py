#!/usr/bin/env python3
"""What a pure write stream of the observation size reaches on this GPU (context for K5's HBM fraction): torch fill_ (128-bit stores) and
cudaMemsetAsync of 2^22 x 311 int64 = 10.4 GB, and a device-to-device copy of the same size (the MEASURED_PEAKS.json pattern)."""
import json


def main():
    import torch

    n = (1 << 22) * 311
    x = torch.empty((n,), dtype=torch.int64, device="cuda")
    y = torch.empty((n // 2,), dtype=torch.int64, device="cuda")

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters / 1e3

    out = {}
    t = timed(lambda: x.fill_(7))
    out["fill_int64_10.4GB"] = {"sec": t, "write_GBps": n * 8 / t / 1e9}
    t = timed(lambda: x.zero_())
    out["zero_10.4GB"] = {"sec": t, "write_GBps": n * 8 / t / 1e9}
    t = timed(lambda: y.copy_(x[: n // 2]))
    out["copy_5.2GB"] = {"sec": t, "read_plus_write_GBps": n * 8 / t / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
