#!/usr/bin/env python3
"""bench.py — full-rules Doppelkopf playout throughput (BASELINE.json metric, configs[1]) plus every other BASELINE config as extra keys.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (C++ restatement) on the host cores

A bench "step" = one pass of the hot path over one batch: 2^24 fresh rs-full-doko games (deal → reservations →
announcements → 48 cards → scoring) played to the end by the random policy WITH announcements
(FdoState::random_action_for_current_player).  metric = game steps (play_action calls) per second, whole job.
Prints ONE JSON line on rank 0.  Extra keys of the line (same run, every N; each its own timed region, W warm-ups, CUDA events, max
over ranks):
    config0_rs_doko_playouts   BASELINE configs[0]: 10^6 rs-doko deals per GPU
    determinizations           BASELINE configs[2]: 65 536 info-states x 4096 card_matching samples per GPU (+ its issue roofline)
    config3_leaf_rollouts      BASELINE configs[3]: 8192 leaves x 1024 rollouts per GPU with dk_allreduce_root_stats (NCCL) INSIDE the timed region
    config4_step_encode        BASELINE configs[4]: 2^22 games lock-step env step + 311-token encode per GPU (HBM roofline)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdo_game_steps_per_sec"
UNIT = "game steps/s"
GAMES_PER_GPU = 1 << 24
SEED = 0xD0C05EED
WORKLOAD = "rs-full-doko full-rules random playouts with announcements, 2^24 concurrent games per B200 (BASELINE configs[1])"
COUNTERS_JSON = os.path.join(ROOT, "profiles", "kernel_counters.json")
K2_NAME = "_ZN2dk24fdo_playout_fresh_kernelILb1EEEvNS_9RngParamsEmPvS2_jPy"
K3_NAME = "_ZN2dk22fdo_determinize_kernelENS_9RngParamsEmjjPK8dk_statePmPhS5_"


def bench_config(n_gpus):
    """The `config` object of the JSON line — identical for the CUDA arm and the reference arm (the reference arm plays a time-bounded
    sample of the same seeded workload per step; its sample size is reported beside `config`, in `cpu_baseline.sample`)."""
    return {"workload": WORKLOAD, "games_per_gpu_per_step": GAMES_PER_GPU, "engine": "rs-full-doko", "policy": "random with announcements",
            "parallelism": f"games sharded over {n_gpus} GPU(s), no collective",
            "l2": "no inputs; 335 MB of outputs per step exceed the 126 MB L2 (extra configs: inputs larger than L2 or an L2 flush between timed iterations, stated per config)"}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            # nvidia-smi's start-up takes driver locks that stall kernel launches for ~0.2 s: wait until it is in its sampling loop
            t0 = time.time()
            while len(self.lines) < 2 and time.time() - t0 < 5.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU baselines (the oracle = C++ restatement of the reference; the Rust toolchain is absent) --------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    L = oracle_lib.load()
    L.orc_cpu_baseline.restype = ctypes.c_double
    L.orc_cpu_baseline.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    return oracle_lib, L


def cpu_baseline(n_threads, target_seconds=12.0):
    """The oracle timed on the host cores over a bounded sample of the same workload: games 0..S-1 of the same Philox stream."""
    oracle_lib, L = _oracle()
    probe = oracle_lib.playout_philox(L, 1, 1 << 14, SEED, 0, 0, True, n_threads)
    rate = (1 << 14) / max(probe["seconds"], 1e-6)
    sample = int(min(max(rate * target_seconds, 1 << 14), 1 << 24))
    r = oracle_lib.playout_philox(L, 1, sample, SEED, 0, 0, True, n_threads)
    steps = int(r["steps"].sum())
    return {"value": steps / r["seconds"], "unit": UNIT, "cores": n_threads if n_threads > 0 else L.orc_hardware_threads(),
            "kind": "port", "sample": f"{sample} games ({steps} game steps) of the same seeded workload, oracle/liboracle.so, "
            f"{r['seconds']:.2f} s", "games_per_sec": sample / r["seconds"]}


def cpu_baseline_kind(kind, units, per_unit, n_threads, unit_name, what):
    """One of the other configs on the host cores (oracle/capi.cpp orc_cpu_baseline): about a second of CPU work each."""
    _, L = _oracle()
    work = ctypes.c_uint64()
    sec = L.orc_cpu_baseline(kind, SEED, units, per_unit, n_threads, ctypes.byref(work))
    return {"value": work.value / sec, "unit": unit_name, "cores": n_threads if n_threads > 0 else L.orc_hardware_threads(), "kind": "port",
            "sample": f"{work.value} {what}, oracle/liboracle.so, {sec:.2f} s"}


def cpu_baseline_config0(n_threads):
    oracle_lib, L = _oracle()
    n = 1_000_000
    r = oracle_lib.playout_philox(L, 0, n, SEED, 0, 0, False, n_threads)
    return {"value": n / r["seconds"], "unit": "games/s", "game_steps_per_sec": float(r["steps"].sum()) / r["seconds"],
            "cores": n_threads if n_threads > 0 else L.orc_hardware_threads(), "kind": "port",
            "sample": f"all {n} rs-doko games of BASELINE configs[0], oracle/liboracle.so, {r['seconds']:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust reference cannot be built in this image
    (no cargo), so this is the C++ restatement (oracle/), all host threads, static partition of game ids — the shape of the
    reference's rayon harness (rs-doko-experiments/src/experiment_0_5.rs:97-147)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    oracle_lib, L = _oracle()
    threads = L.orc_hardware_threads()
    probe = oracle_lib.playout_philox(L, 1, 1 << 14, SEED, 0, 0, True, threads)
    rate = (1 << 14) / max(probe["seconds"], 1e-6)
    budget = 120.0 / max(args.steps + args.warmup, 1)           # whole run within a few minutes
    sample = int(min(max(rate * min(budget, 10.0), 1 << 14), GAMES_PER_GPU))
    for w in range(args.warmup):
        oracle_lib.playout_philox(L, 1, sample, SEED, 0, w, True, threads)
    total_steps, total_sec = 0, 0.0
    for k in range(args.steps):
        r = oracle_lib.playout_philox(L, 1, sample, SEED, 0, 1000 + k, True, threads)
        total_steps += int(r["steps"].sum())
        total_sec += r["seconds"]
    value = total_steps / total_sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "sample_games_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} games per step (a time-bounded sample of the 2^24-game step), {args.steps} steps, oracle/liboracle.so "
                                   "(C++ restatement; Rust toolchain absent)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the other configs on the same host cores (about a second each)
    line["config0_rs_doko_playouts"] = cpu_baseline_config0(threads)
    det = cpu_baseline_kind(3, 4096, 4096, threads, "determinizations/s", "card_matching samples (4096 info-states x 4096)")
    line["determinizations"] = {"value": det["value"], "unit": det["unit"], "cpu_baseline": det,
                                "config": {"workload": "rs-full-doko card_matching, 4096 mid-game info-states x 4096 samples (bounded sample of BASELINE configs[2])"}}
    line["config3_leaf_rollouts"] = cpu_baseline_kind(4, 1024, 1024, threads, "rollouts/s", "determinized leaf rollouts (1024 leaves x 1024)")
    line["config4_step_encode"] = cpu_baseline_kind(5, 1 << 18, 16, threads, "step-encodes/s", "lock-step env steps + encode_state_pi")
    print(json.dumps(line))


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's CPU threads to the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated (first touch puts
    the buffers on that node): with 8 ranks the device→host result copies otherwise all land on whichever node the ranks started on.
    Host-side placement only; returns the node number (-1 when the host exposes none; then nothing is changed)."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local_rank)
        bus = getattr(prop, "pci_bus_id", None)
        dom = getattr(prop, "pci_domain_id", 0)
        devn = getattr(prop, "pci_device_id", 0)
        if bus is None:
            return -1
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return -1
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return -1
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return -1


def load_counters():
    """profiles/kernel_counters.json: ncu counters per kernel, each tied to the sha256 of the SASS it was captured on; compared with the
    hashes of the library that is loaded now (written by the build next to the .so)."""
    from master_doko_reinforcement_learning_b200 import _build

    counters = {}
    if os.path.exists(COUNTERS_JSON):
        counters = json.load(open(COUNTERS_JSON)).get("kernels", {})
    built = _build.sass_hashes()

    def get(mangled):
        c = dict(counters.get(mangled, {}))
        if not c:
            return None
        have, want = built.get(mangled), c.get("sass_sha256")
        c["stale"] = (have != want) if (have and want) else "unknown"
        return c

    return get


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import master_doko_reinforcement_learning_b200 as pkg

    # The JSON line is the ONLY thing this process may print on stdout: libraries that write to file descriptor 1 themselves (NCCL prints
    # its version banner there) are pointed at stderr for the duration of the run; the line goes to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the simulator has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dk = pkg.DokoCuda(local_rank)
    dev = torch.device("cuda", local_rank)
    n = GAMES_PER_GPU
    first_id = rank * n                                 # independent games per rank: no data-path collective
    pts = torch.empty((n, 4), dtype=torch.int32, device=dev)
    steps_out = torch.empty((n,), dtype=torch.int32, device=dev)
    flags = pkg.DK_PLAYOUT_WITH_ANNOUNCEMENTS
    W, K = max(args.warmup, 0), max(args.steps, 1)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_peak_source = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    counter_of = load_counters()

    def one_step(epoch):
        dk.playout(pkg.DK_FDO, n, dk.rng(SEED, first_id, epoch), flags=flags, points_out=pts, steps_out=steps_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    flush_buf = torch.empty((1 << 28,), dtype=torch.uint8, device=dev)      # 256 MiB > the 126 MB L2

    def timed_iters(fn, iters, warm, flush):
        """CUDA-event time of `iters` calls of fn(k) after `warm` untimed ones: (total ms summed over the calls, max over ranks).
        flush=True writes 256 MiB between the calls (outside the event pairs) so that no call starts with its data in L2."""
        for k in range(warm):
            fn(k)
        barrier()
        ms = 0.0
        if flush:
            for k in range(iters):
                flush_buf.fill_(k & 255)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(warm + k)
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(iters):
                fn(warm + k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    # ---- headline: configs[1] ---------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                 # also covers the warm-up: the timed region itself lasts only tens of ms
    for w in range(W):
        one_step(w)
    barrier()
    launches0 = dk.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    total_game_steps = torch.zeros((), dtype=torch.int64, device=dev)
    barrier()
    ev[0].record()
    for k in range(K):
        one_step(1000 + k)
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    launches = dk.launch_count() - launches0       # one playout kernel per step
    # count the work that was done (outside the timed region; deterministic per epoch)
    last_step_game_steps = 0
    for k in range(K):
        one_step(1000 + k)
        last_step_game_steps = steps_out.sum(dtype=torch.int64)
        total_game_steps += last_step_game_steps
    torch.cuda.synchronize()
    last_step_game_steps = int(last_step_game_steps)
    last_point_sum = [int(x) for x in pts.sum(dim=0, dtype=torch.int64).tolist()]
    max_ms = max_over_ranks(elapsed_ms)
    my_steps = int(total_game_steps.item())
    all_steps = sum_over_ranks(my_steps)
    value = all_steps / (max_ms / 1e3)

    # ---- e2e: the host-buffer C-ABI calls; the D2H copy of every step's result is inside the timed region and the result is read on
    # the host.  Four forms of the same work:
    #   summary  dk_playout_summary_host   the batch's statistics (dk_playout_stats, 2160 B per step) — evaluator-style callers
    #   packed   dk_playout_host_packed    per-game results, uint16 points + uint8 steps (3 B per game)
    #   compact  dk_playout_host_compact   per-game results, int8[4] points + uint8 steps (5 B per game)
    #   int32    dk_playout_host           per-game results, int32[4] + uint32 (20 B per game): PCIe bound
    def run_e2e(form):
        bufs = None
        if form == "summary":
            def call(ep):
                return dk.playout_summary_host(pkg.DK_FDO, n, dk.rng(SEED, first_id, ep), flags=flags)
        else:
            dt_p, shape_p, dt_s, fn = {"packed": (torch.int16, (n,), torch.uint8, dk.playout_host_packed),
                                       "compact": (torch.int8, (n, 4), torch.uint8, dk.playout_host_compact),
                                       "int32": (torch.int32, (n, 4), torch.int32, dk.playout_host)}[form]
            bufs = (torch.empty(shape_p, dtype=dt_p).pin_memory(), torch.empty((n,), dtype=dt_s).pin_memory())

            def call(ep):
                fn(pkg.DK_FDO, n, dk.rng(SEED, first_id, ep), flags=flags, points_out=bufs[0], steps_out=bufs[1])
                return None
        call(0)
        barrier()
        t0 = time.perf_counter()
        last = None
        for k in range(K):
            # returns when the step's result is in host memory (the D2H read of the result)
            last = call(1000 + k)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        # same epochs as the counted device pass → same work; verify the last step's host result (outside the timed region)
        if form == "summary":
            assert last.games == n and last.game_steps == last_step_game_steps and [int(x) for x in last.point_sum] == last_point_sum, \
                (last.games, last.game_steps, last_step_game_steps)
        else:
            got = int(bufs[1].sum(dtype=torch.int64))
            assert got == last_step_game_steps, (form, got, last_step_game_steps)
            if form == "packed":
                dec = pkg.api.unpack_points(bufs[0][:1 << 16].numpy().view(np.uint16))
                assert np.array_equal(dec, pts[:1 << 16].cpu().numpy())
        sec = max_over_ranks(sec)
        del bufs
        return all_steps / sec

    e2e = {form: run_e2e(form) for form in ("summary", "packed", "compact", "int32")}

    # ---- BASELINE configs[0]: 10^6 rs-doko games per GPU (the reference's CPU-runnable case) -----------------------------------------
    def run_config0():
        n0 = 1_000_000
        p0 = torch.empty((n0, 4), dtype=torch.int32, device=dev)
        s0 = torch.empty((n0,), dtype=torch.int32, device=dev)
        ms = timed_iters(lambda k: dk.playout(pkg.DK_DOKO, n0, dk.rng(SEED, rank * n0, k), points_out=p0, steps_out=s0), K, max(W, 3), flush=True)
        return {"value": world * K * n0 / (ms / 1e3), "unit": "games/s", "game_steps_per_sec": world * K * n0 * 52 / (ms / 1e3), "ms_per_pass": ms / K,
                "config": {"workload": "rs-doko random-vs-random full-game playouts, 10^6 seeded deals per B200 (BASELINE configs[0])", "games_per_gpu": n0,
                           "l2": "no inputs; 256 MiB L2 flush between the timed iterations (the 20 MB of outputs fit in L2)"}}

    # ---- BASELINE configs[2]: determinizations/s — 65 536 mid-game info-states per GPU (card_index 8 / 16 / 24 / 32 round-robin, observer
    # = seat to move), 4096 consistent hidden-hand samples each (card_matching, rs-full-doko/src/matching/card_matching.rs:31-467),
    # outputs (37 B per sample, 9.9 GB per pass) written to HBM.
    def make_midgame(n_info, epoch0):
        st = dk.new_games(pkg.DK_FDO, n_info, dk.rng(SEED, rank * n_info, 3))
        raw = st.view(torch.uint8).reshape(n_info, 128)
        target = (8 * (1 + (torch.arange(n_info, device=dev) & 3))).to(torch.uint8)
        for it in range(400):                                    # advance every game to its target card index with random legal actions
            active = ((raw[:, 118] < target) & ((raw[:, 124] & 3) != 3)).nonzero().squeeze(1)    # byte 118 = card_index, meta bits 0-1 = phase (3 = finished)
            if active.numel() == 0:
                break
            sub = st[active].contiguous()
            dk.step_random_encode(sub, dk.rng(SEED, rank * n_info, epoch0 + it), want_obs=False)
            st[active] = sub
        return st

    def run_determinizations(passes=3):
        n_info, samples = 1 << 16, 4096
        st = make_midgame(n_info, 100)
        hands = torch.empty((n_info, samples, 4), dtype=torch.int64, device=dev)
        res = torch.empty((n_info, samples, 4), dtype=torch.uint8, device=dev)
        status = torch.empty((n_info, samples), dtype=torch.uint8, device=dev)

        def one_pass(epoch):
            dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_FDO, n_info, samples, pkg.api._ptr(st), ctypes.byref(dk.rng(SEED, rank * n_info, epoch)),
                                          pkg.api._ptr(hands), pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

        ms = timed_iters(one_pass, passes, 3, flush=False)          # 9.9 GB of outputs per pass: nothing survives in L2
        dead = sum_over_ranks(int((status != 0).sum()))
        per_s = world * passes * n_info * samples / (ms / 1e3)
        out = {"value": per_s, "unit": "determinizations/s", "ms_per_pass": ms / passes,
               "config": {"workload": "rs-full-doko card_matching, 65536 mid-game info-states per B200 x 4096 samples (BASELINE configs[2])",
                          "info_states_per_gpu": n_info, "samples_per_info_state": samples, "bytes_written_per_sample": 37,
                          "l2": "9.9 GB of outputs per pass exceed the L2"},
               "dead_ends_last_pass": dead}
        c = counter_of(K3_NAME)
        info = dk.device_info()
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = info["sm_count"] * 4 * sm_mhz * 1e6 / 1e9
        wi = c.get("warp_inst_per_unit") if c else None
        out["roofline"] = {"bound": "issue", "kernel": "fdo_determinize_kernel",
                           "achieved": wi * per_s / world / 1e9 if wi else None, "peak": issue_peak, "unit": "Gwarp-inst/s",
                           "frac": wi * per_s / world / 1e9 / issue_peak if wi else None, "warp_inst_per_sample": wi,
                           "stale": c.get("stale") if c else "no counters", "counters_source": c.get("source") if c else None,
                           "hbm": {"achieved": 37 * per_s / world / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": 37 * per_s / world / 1e9 / hbm_peak,
                                   "algorithmic_bytes_per_sample": 37}}
        # the same configuration for the simplified engine: rs-doko-assignment's sample_assignment (rs-doko-assignment/src/assignment.rs:493-581)
        sd = dk.new_games(pkg.DK_DOKO, n_info, dk.rng(SEED, rank * n_info, 3))
        raw = sd.view(torch.uint8).reshape(n_info, 128)
        target = (8 * (1 + (torch.arange(n_info, device=dev) & 3))).to(torch.uint8)
        for it in range(60):                                      # 4 reservations, then cards up to the target index, with random legal actions
            active = ((raw[:, 118] < target) & ((raw[:, 124] & 3) != 3)).nonzero().squeeze(1)
            if active.numel() == 0:
                break
            sub = sd[active].contiguous()
            act = dk.random_action(pkg.DK_DOKO, sub, dk.rng(SEED, rank * n_info, 300 + it))
            dk.apply(pkg.DK_DOKO, sub, act)
            sd[active] = sub

        def one_pass_doko(epoch):
            dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_DOKO, n_info, samples, pkg.api._ptr(sd), ctypes.byref(dk.rng(SEED, rank * n_info, epoch)),
                                          pkg.api._ptr(hands), pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

        ms_d = timed_iters(one_pass_doko, passes, 3, flush=False)
        out["rs_doko_assignment"] = {"value": world * passes * n_info * samples / (ms_d / 1e3), "unit": "determinizations/s", "ms_per_pass": ms_d / passes,
                                     "dead_ends_last_pass": sum_over_ranks(int((status != 0).sum())),
                                     "config": {"workload": "rs-doko-assignment sample_assignment, 65536 mid-game info-states per B200 x 4096 samples (BASELINE configs[2], simplified engine)"}}
        del hands, res, status
        return out

    # ---- BASELINE configs[3]: leaf-parallel rollouts with the NCCL root-stat reduce.  Every rank holds the same 8192 leaves (config-3
    # states) and plays ITS 1024 rollouts of every leaf (rollout numbers rank*1024 ..), then the exact integer point sums [8192][4] are
    # all-reduced (dk_allreduce_root_stats = ncclAllReduce) so that every rank holds the statistics of all world*1024 rollouts per leaf.
    def run_config3():
        n_leaves, R = 8192, 1024
        leaves_all = make_midgame(n_leaves, 700) if world == 1 else None
        if world > 1:                                              # the same leaves on every rank: rank 0's
            leaves_all = dk.alloc_states(n_leaves)
            if rank == 0:
                leaves_all.copy_(make_midgame(n_leaves, 700))
            dist.broadcast(leaves_all, src=0)
        dk.comm_init()
        sums = torch.zeros((n_leaves, 4), dtype=torch.int64, device=dev)
        ar_ev = []
        out = {}
        for name, det in (("from_determinized_leaves", False), ("determinize_every_rollout", True)):
            ar_ev.clear()

            def one(k, det=det):
                dk.leaf_rollouts(leaves_all, R, dk.rng(SEED, 0, 900 + k, first_sub=rank * R), determinize=det, out=sums)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dk.allreduce_root_stats(sums)
                e1.record()
                ar_ev.append((e0, e1))

            ms = timed_iters(one, K, max(W, 3), flush=True)
            ar_us = [1e3 * a.elapsed_time(b) for a, b in ar_ev[-K:]]
            chk = int((sums[:, 0] * (torch.arange(n_leaves, device=dev) % 1009 + 1)).sum())       # a weighted checksum of seat 0's sums
            same = sum_over_ranks(chk) == world * chk and int(sums.sum(dim=1).abs().max()) == 0    # identical on all ranks, zero-sum
            out[name] = {"value": world * K * n_leaves * R / (ms / 1e3), "unit": "rollouts/s", "ms_per_pass": ms / K,
                         "allreduce_us": {"median": sorted(ar_us)[len(ar_us) // 2], "max": max(ar_us)},
                         "root_stats_identical_on_all_ranks": bool(same)}
        dk.comm_destroy()
        res = out["from_determinized_leaves"]
        res["determinize_every_rollout"] = out["determinize_every_rollout"]
        res["config"] = {"workload": "rs-full-doko leaf-parallel _no_announcement rollouts, 8192 leaves x 1024 rollouts per B200, NCCL all-reduce of the "
                                     "[8192][4] int64 root statistics inside the timed region (BASELINE configs[3])",
                         "leaves": n_leaves, "rollouts_per_leaf_per_gpu": R, "allreduce_bytes": n_leaves * 32,
                         "l2": "1 MB of leaf records; 256 MiB L2 flush between the timed iterations"}
        return res

    # ---- BASELINE configs[4]: lock-step env step + observation encode, 2^22 games per GPU (K5, HBM bound: 128 B read + 128 B written per
    # record + 2488 B of tokens per game = 2744 B) ---------------------------------------------------------------------------------------
    def run_config4():
        n5 = 1 << 22
        st = dk.new_games(pkg.DK_FDO, n5, dk.rng(SEED, rank * n5, 5))
        for k in range(24):
            dk.step_random_encode(st, dk.rng(SEED, rank * n5, k), want_obs=False)
        st0 = st.clone()
        obs = torch.empty((n5, 311), dtype=torch.int64, device=dev)
        act = torch.empty((n5,), dtype=torch.uint8, device=dev)
        ms = timed_iters(lambda k: dk.step_random_encode(st, dk.rng(SEED, rank * n5, 100 + k), obs_out=obs, action_out=act), K, max(W, 3), flush=False)
        # the same steps with narrow observation rows (dk_step_random_encode_narrow: int32 — what the reference's Python side turns the rows
        # into before the network sees them — and uint8), each from the same records
        narrow = {}
        for name, dt, nbytes in (("int32", torch.int32, 4), ("uint8", torch.uint8, 1)):
            stn = st0.clone()
            obn = torch.empty((n5, 311), dtype=dt, device=dev)
            msn = timed_iters(lambda k: dk.step_random_encode_narrow(stn, dk.rng(SEED, rank * n5, 100 + k), obs_out=obn, action_out=act), K, max(W, 3), flush=False)
            per_s_n = world * K * n5 / (msn / 1e3)
            bpg = 256 + 311 * nbytes
            narrow[name] = {"value": per_s_n, "unit": "step-encodes/s", "ms_per_step": msn / K,
                            "roofline": {"bound": "hbm", "kernel": "fdo_step_encode_narrow_kernel", "achieved": bpg * per_s_n / world / 1e9, "peak": hbm_peak,
                                         "unit": "GB/s", "frac": bpg * per_s_n / world / 1e9 / hbm_peak, "algorithmic_bytes_per_game": bpg},
                            "same_records_and_tokens_as_i64": bool(torch.equal(stn.view(torch.uint8), st.view(torch.uint8)) and torch.equal(obn.to(torch.int64), obs))}
            del stn, obn
        del st0
        per_s = world * K * n5 / (ms / 1e3)
        gbs = 2744 * per_s / world / 1e9
        done, _ = dk.terminal(pkg.DK_FDO, st)
        out = {"value": per_s, "unit": "step-encodes/s", "ms_per_step": ms / K,
               "roofline": {"bound": "hbm", "kernel": "fdo_step_encode_tma_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                            "algorithmic_bytes_per_game": 2744, "peak_source": hbm_peak_source},
               "narrow_rows": narrow,
               "games_finished_after_timing": sum_over_ranks(int(done.sum())),
               "config": {"workload": "AlphaZero self-play env step + encode_state_pi (311 x i64), 2^22 games in lock-step per B200 (BASELINE configs[4])",
                          "games_per_gpu": n5, "l2": "512 MB of records in, 10.4 GB of tokens out per step: larger than L2"}}
        del obs, st
        return out

    config0 = run_config0()
    determinizations = run_determinizations()
    config3 = run_config3()
    config4 = run_config4()

    if rank == 0:
        games_per_launch = n
        steps_per_launch = all_steps / max(world * K, 1)
        avg_kernel_s = (sum(kernel_ms) / len(kernel_ms)) / 1e3
        c2 = counter_of(K2_NAME) or {}
        alg_bytes = games_per_launch * 20                       # 16 B points + 4 B step count per game, nothing read
        info = dk.device_info()
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = info["sm_count"] * 4 * sm_mhz * 1e6 / 1e9       # G warp-instructions/s at the clock seen during the run
        wi_per_step = c2.get("warp_inst_per_unit")
        roof = {
            "bound": "issue", "kernel": "fdo_playout_fresh_kernel<true>",
            "achieved": (wi_per_step * steps_per_launch / avg_kernel_s / 1e9) if wi_per_step else None,
            "peak": issue_peak, "unit": "Gwarp-inst/s",
            "frac": (wi_per_step * steps_per_launch / avg_kernel_s / 1e9 / issue_peak) if wi_per_step else None,
            "traffic": c2.get("dram_bytes_per_launch"),
            "peak_source": f"{info['sm_count']} SMs x 4 schedulers x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)",
            "warp_inst_per_game_step": wi_per_step,
            "stale": c2.get("stale", "no counters"),          # False: the counters were captured on exactly the SASS that ran here
            "counters_source": c2.get("source", "profiles/kernel_counters.json missing: run profiles/make_counters.py after the ncu pass"),
            # the pipe that actually limits the kernel: ALU-pipe instructions (logic / select / compare / shift) issue at one warp
            # instruction per TWO cycles per scheduler on sm_100 (measured: profiles/r01_int_rates.json).  ncu's own ALU-pipe
            # utilisation of the captured launch, scaled by (capture duration / duration measured here) for the same number of games.
            "alu_pipe": ({"frac": c2["alu_pipe_pct_of_peak"] / 100.0 * (c2["capture_kernel_ms"] / 1e3 / avg_kernel_s)
                                  * (games_per_launch / c2["capture_games"]),
                          "ncu_pct_at_capture": c2["alu_pipe_pct_of_peak"], "capture_kernel_ms": c2["capture_kernel_ms"],
                          "peak_source": "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active; pipe rate 64 lanes/clk/SM (profiles/r01_int_rates.json)"}
                         if c2.get("alu_pipe_pct_of_peak") and c2.get("capture_kernel_ms") and c2.get("capture_games") else None),
            "hbm": {"bound": "hbm", "achieved": alg_bytes / avg_kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / avg_kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": hbm_peak_source},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": bench_config(world),
            "mean_game_steps_per_game": all_steps / (world * K * n),
            "games_per_sec": world * K * n / (max_ms / 1e3),
            "clocks": clocks,
            "e2e": {"value": e2e["summary"], "unit": UNIT, "h2d_bytes_per_step": 24, "d2h_bytes_per_step": 2160,
                    "api": "dk_playout_summary_host: the step's games are played on the GPU and their statistics (dk_playout_stats: games, game steps, per-seat "
                           "point sums / squares / wins, step histogram) are read into host memory before the call returns",
                    "per_game_forms": {
                        "packed": {"value": e2e["packed"], "d2h_bytes_per_step": n * 3, "api": "dk_playout_host_packed (uint16 points + uint8 steps per game into pinned host buffers)"},
                        "compact": {"value": e2e["compact"], "d2h_bytes_per_step": n * 5, "api": "dk_playout_host_compact (int8[4] + uint8)"},
                        "int32": {"value": e2e["int32"], "d2h_bytes_per_step": n * 20, "api": "dk_playout_host (int32[4] + uint32)",
                                  "pcie_gbs_per_gpu": n * 20 * (e2e["int32"] / world / (all_steps / (world * K))) / 1e9}},
                    "host_numa_node_rank0": numa_node},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(0)
            config0["cpu_baseline"] = cpu_baseline_config0(0)
            determinizations["cpu_baseline"] = cpu_baseline_kind(3, 4096, 4096, 0, "determinizations/s", "card_matching samples (4096 info-states x 4096)")
            determinizations["rs_doko_assignment"]["cpu_baseline"] = cpu_baseline_kind(30, 2048, 512, 0, "determinizations/s", "sample_assignment samples (2048 info-states x 512)")
            config3["cpu_baseline"] = cpu_baseline_kind(4, 1024, 1024, 0, "rollouts/s", "determinized leaf rollouts (1024 leaves x 1024)")
            config4["cpu_baseline"] = cpu_baseline_kind(5, 1 << 18, 16, 0, "step-encodes/s", "lock-step env steps + encode_state_pi")
        line["config0_rs_doko_playouts"] = config0
        line["determinizations"] = determinizations
        line["config3_leaf_rollouts"] = config3
        line["config4_step_encode"] = config4
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
