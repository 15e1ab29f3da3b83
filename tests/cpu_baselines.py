#!/usr/bin/env python3
"""CPU baselines of every config: the C++ oracle — a restatement of the reference, NOT the reference binary (no Rust
toolchain here) — timed on the host cores of the box at 1 thread and at all hardware threads.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
SEED = 0xD0C05EED


def main():
    import oracle_lib

    L = oracle_lib.load()
    L.orc_cpu_baseline.restype = C.c_double
    L.orc_cpu_baseline.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    L.orc_hardware_threads.restype = C.c_int
    hw = L.orc_hardware_threads()
    out = {"kind": "port (oracle/liboracle.so, g++ -O3 -march=native)", "hardware_threads": hw}
    scale = float(os.environ.get("DK_CPU_BASELINE_SCALE", "1"))
    for threads in (1, hw):
        tag = f"{threads}_threads"
        f = threads if threads > 1 else 1
        # configs 1 and 2: whole games
        for name, engine, ann, n in (("config1_doko_playouts", 0, 0, int(40000 * f * scale)), ("config2_fdo_playouts_with_announcements", 1, 1, int(20000 * f * scale)),
                                     ("fdo_playouts_no_announcement_policy", 1, 0, int(20000 * f * scale))):
            import numpy as np

            steps = np.zeros(n, dtype=np.uint32)
            sec = L.orc_playout_philox(engine, ann, SEED, 0, 0, n, threads, None, steps.ctypes.data_as(C.c_void_p), None, None, 0)
            out.setdefault(name, {})[tag] = {"games_per_s": n / sec, "game_steps_per_s": float(steps.sum()) / sec, "sec": sec}
        for name, kind, units, per in (("config3_card_matching_samples", 3, int(64 * f * scale), 512), ("config3_sample_assignment_samples", 30, int(64 * f * scale), 256),
                                       ("config4_determinized_leaf_rollouts", 4, int(64 * f * scale), 128), ("config5_env_step_plus_encode_pi", 5, int(4000 * f * scale), 16),
                                       ("n2_flat_mc_rollouts", 7, int(64 * f * scale), 64), ("n3_uct_iterations", 6, int(32 * f * scale), 512)):
            work = C.c_uint64()
            sec = L.orc_cpu_baseline(kind, SEED, units, per, threads, C.byref(work))
            out.setdefault(name, {})[tag] = {"items_per_s": work.value / sec, "items": work.value, "sec": sec}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
