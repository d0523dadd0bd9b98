#!/usr/bin/env python
"""Benchmark of the Yacht-Auction hot path (BASELINE.json metric: game steps/s, MCTS sims/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): batched stepping + legal-move enumeration, 65,536 concurrent
games per GPU, uniform random legal policy, legal mask (uint8[3226], the reference's dtype)
materialised every ply.  One bench "step" = 48 fused plies = one full 13-round game for every
game slot (finished games are re-dealt on device), i.e. 48 kernel launches.
Games are independent: under torchrun every rank owns `games` game slots (weak scaling, no
collective on the data path); global game ids are rank * games + g.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PLIES_PER_GAME = 48
ACTION_SIZE = 3226
ALGO_BYTES_PER_STEP = 32 + 32 + 4 + ACTION_SIZE       # 3,294 B per game-step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="concurrent games per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernels / MCTS lines")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": "configs[1]: batched stepping + legal-move enumeration, %d concurrent games per GPU, "
                    "uniform random legal policy, uint8[3226] legal mask materialised every ply" % args.games,
        "games_per_gpu": args.games, "plies_per_step": PLIES_PER_GAME, "mask": "uint8[3226] per game per ply",
        "sharding": "game-index x%d (no collective on the path)" % n_gpus,
        "l2": "211 MB of mask output per launch > 126 MB L2 (no explicit flush needed)",
        "launch": "48 plies captured once as a CUDA graph, one replay per step",
    }


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import arena_port
    total = args.steps + args.warmup
    budget = max(0.5, min(4.0, 150.0 / max(total, 1)))
    for _ in range(args.warmup):
        arena_port.timed_sample(seed=args.seed, budget_s=min(budget, 1.0))
    t0 = time.perf_counter()
    steps_done, rate_sum = 0, 0.0
    cores = os.cpu_count() or 1
    for i in range(args.steps):
        r = arena_port.timed_sample(seed=args.seed + 1 + i, budget_s=budget)
        steps_done += r["steps"]
        rate_sum += r["steps_per_s"]
        cores = r["cores"]
    wall = time.perf_counter() - t0
    value = rate_sum / max(args.steps, 1)
    sample = "%d bounded samples of ~%.1f s: Arena-style random-vs-random full games (oracle port of " \
             "Arena.playGame + RandomYachtPlayer), %d processes" % (args.steps, budget, cores)
    line = {
        "impl": "reference", "metric": "game_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "game_steps_played": steps_done,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)

    def summary(self):
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v for v in vis.split(",") if v != ""]
        if local < len(ids) and ids[local].isdigit():
            return int(ids[local])
    return local


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nypc_yacht_auction_b200 import _lib
    from nypc_yacht_auction_b200.engine import BatchedYacht

    n = args.games
    env = BatchedYacht(n, seed=args.seed, game_base=rank * n, device=dev)
    masks = torch.empty((n, ACTION_SIZE), dtype=torch.uint8, device=dev)
    wins = torch.zeros(3, dtype=torch.int64, device=dev)

    game_graph = env.capture_game_graph(masks=masks, plies=PLIES_PER_GAME, auto_reset=True)

    def one_step():                      # 48 fused plies = 48 kernel launches, replayed as one CUDA graph
        game_graph.replay()

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        one_step()
    fence()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    t_wall = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    fence()
    ms = ev0.elapsed_time(ev1)
    wall = time.perf_counter() - t_wall
    clocks_note = "sampled during the timed region"
    if len(sampler.rows) < 3:
        # timed region shorter than the sampling period: keep the same kernel running ~1.5 s and sample that
        clocks_note = "timed region < sampler period; sampled over 1.5 s of the same launches right after it"
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            one_step()
            torch.cuda.synchronize(dev)
    sampler.stop()
    assert int(env.err_flag.item()) == 0, "engine reported a rule error"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = args.steps * PLIES_PER_GAME
    game_steps = launches * n * world
    value = game_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (ya_k_play_ply): algorithmic bytes / mean launch time
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    launch_ms = ms / launches
    achieved = ALGO_BYTES_PER_STEP * n / (launch_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("ya_k_play_ply", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "ya_k_play_ply", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STEP * n, "launch_us": launch_ms * 1e3}

    # ---- end to end through the host-buffer C ABI (pinned host buffers, copies inside the timed region)
    e2e = run_e2e(args, torch, dist, _lib, dev, rank, world, n)

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, torch, env, dev, n, peak, rank, world, dist)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import arena_port
        r = arena_port.timed_sample(seed=args.seed, budget_s=args.cpu_seconds)
        cpu = {"value": r["steps_per_s"], "unit": "steps/s", "cores": r["cores"], "kind": "port",
               "sample": "%d full random-vs-random games (%d plies) in %.1f s on %d processes: oracle port of "
                         "Arena.playGame + RandomYachtPlayer" % (r["games"], r["steps"], r["seconds"], r["cores"])}
        if not args.no_extras:                # CPU arm of the MCTS lines, bounded sample
            m = arena_port.timed_mcts_sample(seed=args.seed, sims=25, budget_s=min(8.0, args.cpu_seconds))
            cpu["mcts_uniform_sims_per_s"] = m["sims_per_s"]
            cpu["mcts_sample"] = "oracle port of MCTS.py self-play, uniform evaluator, numMCTSSims=25, %d sims in %.1f s on %d processes" % (
                m["sims"], m["seconds"], m["cores"])
            m = arena_port.timed_mcts_nn_sample(seed=args.seed, sims=100, budget_s=min(6.0, args.cpu_seconds))
            cpu["mcts_nn_sims_per_s"] = m["sims_per_s"]
            cpu["mcts_nn_sample"] = ("oracle port of MCTS.py + a float32 CPU forward of the same network per leaf (batch 1, as "
                                     "NNetWrapper.predict), numMCTSSims=100, %d sims in %.1f s on %d processes" % (
                                         m["sims"], m["seconds"], m["cores"]))
        try:                                  # context only: the same workload in plain C (OpenMP), not the reference's cost
            from oracle import c_oracle
            t0 = time.perf_counter()
            c_steps = c_oracle.timed_steps(400000, PLIES_PER_GAME, args.seed)
            cpu["c_oracle_steps_per_s"] = c_steps / (time.perf_counter() - t0)
        except Exception as exc:              # noqa: BLE001 -- the C oracle is optional test infrastructure
            cpu["c_oracle_steps_per_s"] = None
            cpu["c_oracle_error"] = str(exc)[:100]

    if rank == 0:
        clocks = sampler.summary()
        clocks["note"] = clocks_note
        line = {
            "metric": "game_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, torch, dist, _lib, dev, rank, world, n):
    """The same 48-ply step through the host-buffer C ABI: boards live in pinned HOST memory as 64-byte
    records; every bench step (= 48 plies = one full game per slot, like one Arena.playGames call) copies
    them host->device, runs the 48 fused plies (mask materialised in HBM, exactly like the device-resident
    path) and copies the records (boards, outcomes, win tallies) back.  The variant with a host round trip
    after EVERY ply is reported next to it."""
    import ctypes
    lib = _lib.load()
    handle = ctypes.c_void_p()
    _lib.check(lib.ya_host_create(n, 1, ctypes.byref(handle)), "ya_host_create")
    rec = torch.zeros((n, 16), dtype=torch.int32, pin_memory=True)
    h_err = torch.zeros(1, dtype=torch.int32, pin_memory=True)
    from nypc_yacht_auction_b200.engine import BatchedYacht
    tmp = BatchedYacht(n, seed=args.seed + 17, game_base=rank * n, device=dev)
    st = tmp.states.cpu()                                    # [2, n, 4] planes -> words 0..7 of each record
    rec[:, 0:4] = st[0]
    rec[:, 4:8] = st[1]
    rec[:, 10] = 1                                           # player
    del tmp

    def host_step():            # ONE host call = one bench step: 48 plies = one full game for every slot
        _lib.check(lib.ya_host_play_plies_records(handle, _lib.ptr(rec), PLIES_PER_GAME, None, _lib.ptr(h_err),
                                                  args.seed + 17, rank * n, 1), "ya_host_play_plies_records")

    def host_step_per_ply():    # host round trip after every ply (what a host-side policy would need)
        for _ in range(PLIES_PER_GAME):
            _lib.check(lib.ya_host_play_ply_records(handle, _lib.ptr(rec), None, _lib.ptr(h_err), args.seed + 17,
                                                    rank * n, 1), "ya_host_play_ply_records")

    def timed(fn, steps):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()                             # synchronous: returns after the D2H copies completed
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    steps = max(1, min(args.steps, 50))
    dt = timed(host_step, steps)
    finished = int(rec[:, 13].min())
    assert finished >= steps, "every slot must have finished a game per step"
    steps_pp = max(1, min(args.steps, 10))
    dt_pp = timed(host_step_per_ply, steps_pp)
    assert int(h_err.item()) == 0
    lib.ya_host_destroy(handle)
    return {"value": steps * PLIES_PER_GAME * n * world / dt, "unit": "steps/s",
            "h2d_bytes_per_step": n * 64, "d2h_bytes_per_step": n * 64 + 4,
            "steps": steps, "api": "ya_host_play_plies_records(plies=48) (C ABI): per bench step the 64-byte game records go "
            "host->device from pinned memory, 48 fused plies run (uint8 mask materialised in HBM every ply, finished games "
            "re-dealt and tallied), the records (boards, last actions, outcomes, win tallies) come back; 4 slices pipelined "
            "over 4 streams",
            "per_ply_round_trip": {"value": steps_pp * PLIES_PER_GAME * n * world / dt_pp, "unit": "steps/s",
                                   "h2d_bytes_per_step": n * 64 * PLIES_PER_GAME, "d2h_bytes_per_step": (n * 64 + 4) * PLIES_PER_GAME,
                                   "api": "ya_host_play_ply_records: host round trip after every ply (PCIe-bound)"},
            "timing": "host wall clock around synchronous calls, max over ranks"}


def run_extras(args, torch, env, dev, n, peak, rank=0, world=1, dist=None):
    """Secondary kernels of the path, timed alone (CUDA events, 3 warm-ups)."""
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    # bring every game to a score ply with 10 dice (ply 4) so the enumeration does full work
    env.episode.zero_()
    env.reset()
    for _ in range(4):
        env.play_ply(masks=None, auto_reset=False)
    table = torch.empty((n, 12, 252), dtype=torch.uint8, device=dev)
    ms = timed(lambda: env.enumerate_scores(out=table), 50)
    bytes_ = n * (32 + 3024)
    out["enumerate_scores"] = {"score_plies_per_sec": n / (ms * 1e-3), "us_per_launch": ms * 1e3,
                               "achieved_gbs": bytes_ / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peak,
                               "algorithmic_bytes_per_game": 32 + 3024}
    ms = timed(lambda: env.play_ply(masks=None, auto_reset=True), 200)
    out["transition_only"] = {"steps_per_sec": n / (ms * 1e-3), "us_per_launch": ms * 1e3,
                              "note": "fused ply without materialising the mask (68 B/step; L2-resident, latency-bound)"}
    from nypc_yacht_auction_b200 import mcts_bench
    out.update(mcts_bench.run(args, torch, dev, rank, world, dist))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if "nn_forward" in out and os.path.exists(peaks_path):          # tensor-bound kernel: against the measured bf16 GEMM rate
        tf = json.load(open(peaks_path)).get("bf16_tflops")
        if tf:
            out["nn_forward"].update({"bound": "tensor", "peak_tflops": float(tf), "frac": out["nn_forward"]["tflops"] / float(tf),
                                      "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS 8192^3 burst)"})
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
