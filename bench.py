#!/usr/bin/env python3
"""bench.py — full-rules Doppelkopf playout throughput (BASELINE.json metric, config[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (C++ restatement) on the host cores

A bench "step" = one pass of the hot path over one batch: 2^24 fresh rs-full-doko games (deal → reservations →
announcements → 48 cards → scoring) played to the end by the random policy WITH announcements
(FdoState::random_action_for_current_player).  metric = game steps (play_action calls) per second, whole job.
Prints ONE JSON line on rank 0.
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
PROFILE_JSON = os.path.join(ROOT, "profiles", "fdo_playout_counters.json")


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


def cpu_baseline(n_threads, target_seconds=12.0):
    """The oracle (C++ restatement of the reference's rules) timed on the host cores over a bounded sample of the same
    workload: games 0..S-1 of the same Philox stream."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    L = oracle_lib.load()
    probe = oracle_lib.playout_philox(L, 1, 1 << 14, SEED, 0, 0, True, n_threads)
    rate = (1 << 14) / max(probe["seconds"], 1e-6)
    sample = int(min(max(rate * target_seconds, 1 << 14), 1 << 24))
    r = oracle_lib.playout_philox(L, 1, sample, SEED, 0, 0, True, n_threads)
    steps = int(r["steps"].sum())
    return {"value": steps / r["seconds"], "unit": UNIT, "cores": n_threads if n_threads > 0 else L.orc_hardware_threads(),
            "kind": "port", "sample": f"{sample} games ({steps} game steps) of the same seeded workload, oracle/liboracle.so, "
            f"{r['seconds']:.2f} s", "games_per_sec": sample / r["seconds"]}


def cpu_baseline_determinizations(n_threads):
    """card_matching of the oracle on the host cores: 4096 info-states made like BASELINE configs[2] x 4096 samples."""
    import ctypes as C

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    L = oracle_lib.load()
    L.orc_cpu_baseline.restype = C.c_double
    L.orc_cpu_baseline.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    work = C.c_uint64()
    sec = L.orc_cpu_baseline(3, SEED, 4096, 4096, n_threads, C.byref(work))
    return {"value": work.value / sec, "unit": "determinizations/s", "cores": n_threads if n_threads > 0 else L.orc_hardware_threads(), "kind": "port",
            "sample": f"{work.value} samples (4096 info-states x 4096), oracle/liboracle.so, {sec:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust reference cannot be built in this image
    (no cargo), so this is the C++ restatement (oracle/), all host threads, static partition of game ids — the shape of the
    reference's rayon harness (rs-doko-experiments/src/experiment_0_5.rs:97-147)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    L = oracle_lib.load()
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
        "config": {"workload": WORKLOAD, "sample_games_per_step": sample, "engine": "rs-full-doko", "policy": "random with announcements"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} games per step, {args.steps} steps, oracle/liboracle.so (C++ restatement; Rust toolchain absent)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    det = cpu_baseline_determinizations(threads)                # the other half of the metric on the same host cores (about 1.3 s)
    line["determinizations"] = {"value": det["value"], "unit": det["unit"], "cpu_baseline": det,
                                "config": {"workload": "rs-full-doko card_matching, 4096 mid-game info-states x 4096 samples (bounded sample of BASELINE configs[2])"}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's CPU threads to the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated (first touch puts
    the buffers on that node): with 8 ranks the device→host result copies otherwise all land on whichever node the ranks started on.
    Host-side placement only; returns the node number or None when the topology cannot be read (then nothing is changed)."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        devn = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        if bus is None:
            return None
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import master_doko_reinforcement_learning_b200 as pkg

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the simulator has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
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

    def one_step(epoch):
        dk.playout(pkg.DK_FDO, n, dk.rng(SEED, first_id, epoch), flags=flags, points_out=pts, steps_out=steps_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                 # also covers the warm-up: the timed region itself lasts only tens of ms
    for w in range(max(args.warmup, 0)):
        one_step(w)
    barrier()
    launches0 = dk.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    total_game_steps = torch.zeros((), dtype=torch.int64, device=dev)
    barrier()
    ev[0].record()
    for k in range(args.steps):
        one_step(1000 + k)
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    launches = dk.launch_count() - launches0       # one playout kernel per step
    # count the work that was done (outside the timed region; deterministic per epoch)
    last_step_game_steps = 0
    for k in range(args.steps):
        one_step(1000 + k)
        last_step_game_steps = steps_out.sum(dtype=torch.int64)
        total_game_steps += last_step_game_steps
    torch.cuda.synchronize()
    last_step_game_steps = int(last_step_game_steps)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    tot = total_game_steps.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms = float(t.item())
    all_steps = int(tot.item())
    value = all_steps / (max_ms / 1e3)

    # ---- e2e: the host-buffer C-ABI calls (pinned host outputs; the D2H copies are inside the timed region and the step's result —
    # every game's points and step count — is read on the host).  Headline e2e = dk_playout_host_compact (int8 points + uint8 steps,
    # lossless, 5 B/game over PCIe); the int32/uint32 form (20 B/game) is reported beside it.
    def run_e2e(compact):
        if compact:
            h_pts = torch.empty((n, 4), dtype=torch.int8).pin_memory()
            h_steps = torch.empty((n,), dtype=torch.uint8).pin_memory()
            call = dk.playout_host_compact
        else:
            h_pts = torch.empty((n, 4), dtype=torch.int32).pin_memory()
            h_steps = torch.empty((n,), dtype=torch.int32).pin_memory()
            call = dk.playout_host
        call(pkg.DK_FDO, n, dk.rng(SEED, first_id, 0), flags=flags, points_out=h_pts, steps_out=h_steps)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            # returns when every game's points and step count of this step are in host memory (the D2H read of the result)
            call(pkg.DK_FDO, n, dk.rng(SEED, first_id, 1000 + k), flags=flags, points_out=h_pts, steps_out=h_steps)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        # same epochs as the counted device pass → same work; verify on the last step's host result (outside the timed region)
        last = int(h_steps.sum(dtype=torch.int64))
        assert last == last_step_game_steps, (last, last_step_game_steps)
        game_steps = int(total_game_steps.item())
        te = torch.tensor([sec], dtype=torch.float64, device=dev)
        se = torch.tensor([game_steps], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(se, op=dist.ReduceOp.SUM)
        del h_pts, h_steps
        return int(se.item()) / float(te.item())

    e2e_value = run_e2e(True)
    e2e_int32 = run_e2e(False)

    # ---- second half of BASELINE.json's metric: determinizations/s on BASELINE configs[2] — 65 536 mid-game info-states per GPU
    # (card_index 8 / 16 / 24 / 32 round-robin, observer = seat to move), 4096 consistent hidden-hand samples each (card_matching,
    # rs-full-doko/src/matching/card_matching.rs:31-467), outputs (37 B per sample, 9.9 GB per pass) written to HBM.
    def run_determinizations(passes=3):
        n_info, samples = 1 << 16, 4096
        st = dk.new_games(pkg.DK_FDO, n_info, dk.rng(SEED, rank * n_info, 3))
        raw = st.view(torch.uint8).reshape(n_info, 128)
        target = (8 * (1 + (torch.arange(n_info, device=dev) & 3))).to(torch.uint8)
        for it in range(400):                                    # advance every game to its target card index with random legal actions
            active = ((raw[:, 118] < target) & ((raw[:, 124] & 3) != 3)).nonzero().squeeze(1)    # byte 118 = card_index, meta bits 0-1 = phase (3 = finished)
            if active.numel() == 0:
                break
            sub = st[active].contiguous()
            dk.step_random_encode(sub, dk.rng(SEED, rank * n_info, 100 + it), want_obs=False)
            st[active] = sub
        hands = torch.empty((n_info, samples, 4), dtype=torch.int64, device=dev)
        res = torch.empty((n_info, samples, 4), dtype=torch.uint8, device=dev)
        status = torch.empty((n_info, samples), dtype=torch.uint8, device=dev)

        def one_pass(epoch):
            dk._check(dk.L.dk_determinize(dk.ctx, pkg.DK_FDO, n_info, samples, pkg.api._ptr(st), ctypes.byref(dk.rng(SEED, rank * n_info, epoch)),
                                          pkg.api._ptr(hands), pkg.api._ptr(res), pkg.api._ptr(status), dk._stream()), "dk_determinize")

        one_pass(0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(passes):
            one_pass(1 + k)
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dead = (status != 0).sum(dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dead, op=dist.ReduceOp.SUM)
        ms = float(tt.item())
        out = {"value": world * passes * n_info * samples / (ms / 1e3), "unit": "determinizations/s", "ms_per_pass": ms / passes,
               "config": {"workload": "rs-full-doko card_matching, 65536 mid-game info-states per B200 x 4096 samples (BASELINE configs[2])",
                          "info_states_per_gpu": n_info, "samples_per_info_state": samples, "bytes_written_per_sample": 37},
               "dead_ends_last_pass": int(dead.item())}
        del hands, res, status
        return out

    determinizations = run_determinizations()

    if rank == 0:
        games_per_launch = n
        steps_per_launch = all_steps / max(world * args.steps, 1)
        avg_kernel_s = (sum(kernel_ms) / len(kernel_ms)) / 1e3
        counters = {}
        if os.path.exists(PROFILE_JSON):
            counters = json.load(open(PROFILE_JSON))
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = games_per_launch * 20                       # 16 B points + 4 B step count per game, nothing read
        info = dk.device_info()
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = info["sm_count"] * 4 * sm_mhz * 1e6 / 1e9       # G warp-instructions/s at the clock seen during the run
        wi_per_step = counters.get("warp_inst_per_game_step")
        roof = {
            "bound": "issue", "kernel": "fdo_playout_fresh_kernel<true>",
            "achieved": (wi_per_step * steps_per_launch / avg_kernel_s / 1e9) if wi_per_step else None,
            "peak": issue_peak, "unit": "Gwarp-inst/s",
            "frac": (wi_per_step * steps_per_launch / avg_kernel_s / 1e9 / issue_peak) if wi_per_step else None,
            "traffic": counters.get("dram_bytes_per_launch"),
            "peak_source": f"{info['sm_count']} SMs x 4 schedulers x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)",
            "warp_inst_per_game_step": wi_per_step,
            "counters_source": counters.get("source", "profiles/fdo_playout_counters.json missing: run the ncu pass"),
            # the pipe that actually limits the kernel: ALU-pipe instructions (logic / select / compare / shift) issue at one warp
            # instruction per TWO cycles per scheduler on sm_100 (measured: profiles/r01_int_rates.json).  ncu's own ALU-pipe
            # utilisation of the captured launch, scaled by (capture duration / duration measured here) for the same number of games.
            "alu_pipe": ({"frac": counters["alu_pipe_pct_of_peak"] / 100.0 * (counters["capture_kernel_ms"] / 1e3 / avg_kernel_s)
                                  * (games_per_launch / counters["capture_games"]),
                          "ncu_pct_at_capture": counters["alu_pipe_pct_of_peak"], "capture_kernel_ms": counters["capture_kernel_ms"],
                          "peak_source": "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active; pipe rate 64 lanes/clk/SM (profiles/r01_int_rates.json)"}
                         if counters.get("alu_pipe_pct_of_peak") and counters.get("capture_kernel_ms") else None),
            "hbm": {"bound": "hbm", "achieved": alg_bytes / avg_kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / avg_kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"},
        }
        base = cpu_baseline(0) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "games_per_gpu_per_step": n, "engine": "rs-full-doko", "policy": "random with announcements",
                       "mean_game_steps_per_game": all_steps / (world * args.steps * n), "parallelism": f"games sharded over {world} GPU(s), no collective",
                       "l2": "no inputs; 335 MB of outputs per step exceed the 126 MB L2"},
            "games_per_sec": world * args.steps * n / (max_ms / 1e3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24, "d2h_bytes_per_step": n * 5,
                    "api": "dk_playout_host_compact (int8 points + uint8 steps into pinned host buffers, chunked so copies overlap the kernel)",
                    "int32_api": {"value": e2e_int32, "d2h_bytes_per_step": n * 20, "api": "dk_playout_host"},
                    "host_numa_node_rank0": numa_node},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if base:
            line["cpu_baseline"] = base
            determinizations["cpu_baseline"] = cpu_baseline_determinizations(0)
        line["determinizations"] = determinizations
        print(json.dumps(line))
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
