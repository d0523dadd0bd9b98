#!/usr/bin/env python
"""Benchmark of the Yacht-Auction hot path (BASELINE.json metric: game steps/s AND MCTS simulations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1]): batched stepping + legal-move enumeration, 65,536 concurrent games per GPU,
uniform random legal policy, legal mask (uint8[3226], the reference's dtype) materialised every ply.  One bench "step" =
48 fused plies = one full 13-round game for every game slot (finished games are re-dealt on device) = 48 kernel launches.
Games are independent: under torchrun every rank owns `games` slots (weak scaling, no collective on the data path);
global game ids are rank * games + g.

The same JSON line carries first-class blocks for the MCTS half of the metric, each with its own roofline, cpu_baseline
and e2e (full 48-ply episodes): `mcts_uniform` (configs[2]), `mcts_nn` (configs[3]) and `mcts_nn_1m` (configs[4]:
1,048,576 games split over the ranks, strong scaling).  See DESIGN.md "Measurement" for every field.

Prints ONE JSON line (rank 0).
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
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernels and every MCTS block")
    ap.add_argument("--no-1m", action="store_true", help="skip the configs[4] block (1,048,576 games, ~1 min on one GPU)")
    ap.add_argument("--games-1m", type=int, default=1 << 20, help="total games of the configs[4] block (all ranks together)")
    ap.add_argument("--wave-1m", type=int, default=2 * 148 * 128, help="games per wave and tree pool of the configs[4] block")
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


# ----------------------------------------------------------------------------- the real reference on the host CPU
def reference_dir():
    """Where the UNMODIFIED reference can be imported from: baseline/_ref (travels to the GPU box; baseline/install_ref.py)
    or /root/reference (build container only)."""
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(d, "yacht", "YachtGame.py")):
            return d
    return None


def _ref_proc(ref, *argv):
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    return subprocess.Popen([sys.executable, os.path.join(ROOT, "baseline", "run_reference.py"), ref, *[str(a) for a in argv]],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, cwd=ROOT)


def _ref_result(proc, timeout=600):
    try:
        out, _ = proc.communicate(timeout=timeout)
        return json.loads(out.strip().splitlines()[-1])
    except Exception as exc:              # noqa: BLE001 -- a reported baseline must not take the bench down
        proc.kill()
        return {"error": str(exc)[:120]}


def time_reference_python(arena_games=1000, mcts_games=3, nn_games=1):
    """BASELINE.md section 4 / SURVEY.md 8(d) config 1, exactly: the reference's own Arena + RandomYachtPlayer for 1,000 games,
    its MCTS with the uniform evaluator (25 sims, 3 games) and with NNetWrapper(cuda=False) (25 sims, 1 game): one Python
    process = one core each, the three run side by side on three cores."""
    ref = reference_dir()
    if ref is None:
        return {"unavailable": "no baseline/_ref and no /root/reference on this machine"}
    procs = {"arena": _ref_proc(ref, "arena", arena_games), "mcts_uniform": _ref_proc(ref, "mcts", 25, mcts_games),
             "mcts_nn": _ref_proc(ref, "mcts_nn", 25, nn_games)}
    res = {k: _ref_result(p) for k, p in procs.items()}
    out = {"cores": 1, "source": os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref,
           "what": "the unmodified reference through its own API, one single-threaded Python process per figure"}
    a, m, n = res["arena"], res["mcts_uniform"], res["mcts_nn"]
    if "steps_per_s" in a:
        out.update({"steps_per_s": a["steps_per_s"], "games_per_s": a["games_per_s"], "arena_sample": "%s: %d games in %.1f s" % (a["api"], a["games"], a["seconds"])})
    if "sims_per_s" in m:
        out.update({"mcts_uniform_sims_per_s": m["sims_per_s"], "mcts_uniform_sample": "%s: %d games, %d sims in %.1f s" % (m["api"], m["games"], m["sims"], m["seconds"])})
    if "sims_per_s" in n:
        out.update({"mcts_nn_sims_per_s": n["sims_per_s"], "mcts_nn_sample": "%s: %d games, %d sims in %.1f s" % (n["api"], n["games"], n["sims"], n["seconds"])})
    errs = {k: v["error"] for k, v in res.items() if "error" in v}
    if errs:
        out["errors"] = errs
    return out


def reference_all_cores(seconds, seed):
    """The reference's Arena random-vs-random on every host core: one unmodified single-threaded process per core."""
    ref = reference_dir()
    cores = os.cpu_count() or 1
    procs = [_ref_proc(ref, "arena_for", seconds, seed * 1000 + i) for i in range(cores)]
    res = [_ref_result(p) for p in procs]
    good = [r for r in res if "steps_per_s" in r]
    return {"steps_per_s": sum(r["steps_per_s"] for r in good), "steps": sum(r["steps"] for r in good), "games": sum(r["games"] for r in good),
            "cores": len(good), "seconds": max([r["seconds"] for r in good] or [0.0])}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    budget = max(0.5, min(4.0, 150.0 / max(total, 1)))
    ref = reference_dir()
    if ref is not None:
        kind = "reference"
        sample_one = lambda seed, b: reference_all_cores(b, seed)
        what = "the UNMODIFIED reference (%s): Arena(RandomYachtPlayer, RandomYachtPlayer).playGames on every host core, one " \
               "single-threaded Python process per core" % (os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref)
    else:
        from oracle import arena_port
        kind = "port"
        sample_one = lambda seed, b: arena_port.timed_sample(seed=seed, budget_s=b)
        what = "oracle port of Arena.playGame + RandomYachtPlayer (the reference is not on this machine), one process per core"
    for _ in range(args.warmup):
        sample_one(args.seed, min(budget, 1.0))
    t0 = time.perf_counter()
    steps_done, rate_sum = 0, 0.0
    cores = os.cpu_count() or 1
    for i in range(args.steps):
        r = sample_one(args.seed + 1 + i, budget)
        steps_done += r["steps"]
        rate_sum += r["steps_per_s"]
        cores = r["cores"]
    wall = time.perf_counter() - t0
    value = rate_sum / max(args.steps, 1)
    sample = "%d bounded samples of ~%.1f s of full random-vs-random games: %s, %d processes" % (args.steps, budget, what, cores)
    line = {
        "impl": "reference", "metric": "game_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "game_steps_played": steps_done,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.rows = []
        self.proc = None
        self.index = index
        self.period_ms = period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "tensor": float(d["bf16_tflops"]), "tensor_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "MEASURED_PEAKS.json (measured copy / cuBLAS bf16 burst)"}
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_summary():
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(path))
    except Exception:                     # noqa: BLE001
        return {}


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

    peaks = load_peaks()
    prof = ncu_summary()
    n = args.games
    env = BatchedYacht(n, seed=args.seed, game_base=rank * n, device=dev)
    masks = torch.empty((n, ACTION_SIZE), dtype=torch.uint8, device=dev)

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
    sampler.stop()
    clocks = sampler.summary()
    clocks["note"] = "sampled during the timed region" if clocks["samples"] >= 3 else \
        "timed region shorter than three sampler periods: see `sustained` for clocks under the same load over >= 2 s"
    # ---- the same launches for >= 2 s (independent of --steps), clocks sampled INSIDE the region
    sus_sampler = ClockSampler(physical_gpu_index(local))
    per_step_ms = max(ms / max(args.steps, 1), 1e-3)
    sus_steps = int(2500.0 / per_step_ms) + 1
    fence()
    sus_sampler.start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(sus_steps):
        one_step()
    s1.record()
    fence()
    sus_ms = s0.elapsed_time(s1)
    sus_sampler.stop()
    assert int(env.err_flag.item()) == 0, "engine reported a rule error"
    if world > 1:
        t = torch.tensor([ms, sus_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, sus_ms = float(t[0].item()), float(t[1].item())
    launches = args.steps * PLIES_PER_GAME
    game_steps = launches * n * world
    value = game_steps / (ms * 1e-3)
    sus_clocks = sus_sampler.summary()
    if clocks["samples"] < 3:                      # the contract's clocks record must come from load: use the sustained samples
        clocks.update({k: sus_clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")})
    sustained = {"seconds": sus_ms * 1e-3, "steps": sus_steps, "value": sus_steps * PLIES_PER_GAME * n * world / (sus_ms * 1e-3),
                 "unit": "steps/s", "clocks": sus_clocks,
                 "roofline_frac": ALGO_BYTES_PER_STEP * n / (sus_ms * 1e-3 / (sus_steps * PLIES_PER_GAME)) / 1e9 / peaks["hbm"]}

    # ---- roofline of the dominant kernel (ya_k_play_ply): algorithmic bytes / mean launch time
    launch_ms = ms / launches
    achieved = ALGO_BYTES_PER_STEP * n / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "ya_k_play_ply", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": achieved / peaks["hbm"], "traffic": prof.get("ya_k_play_ply", {}).get("dram_bytes_per_launch"),
                "peak_source": peaks["source"], "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STEP * n, "launch_us": launch_ms * 1e3}

    # ---- end to end through the host-buffer C ABI (pinned host buffers, copies inside the timed region)
    e2e = run_e2e(args, torch, dist, _lib, dev, rank, world, n)
    del masks, game_graph

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, torch, env, dev, n, peaks, prof, rank, world, dist)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = run_cpu_baselines(args, extras, world)

    if rank == 0:
        line = {
            "metric": "game_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall, "sustained": sustained,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_cpu_baselines(args, extras, world):
    """Host-CPU figures, rank 0, after every GPU measurement: the oracle port on all cores (bounded samples) and -- when the
    reference itself is on this machine (baseline/_ref) -- the UNMODIFIED reference, config 1 exactly."""
    from oracle import arena_port
    r = arena_port.timed_sample(seed=args.seed, budget_s=args.cpu_seconds)
    cpu = {"value": r["steps_per_s"], "unit": "steps/s", "cores": r["cores"], "kind": "port",
           "sample": "%d full random-vs-random games (%d plies) in %.1f s on %d processes: oracle port of "
                     "Arena.playGame + RandomYachtPlayer" % (r["games"], r["steps"], r["seconds"], r["cores"])}
    ref = time_reference_python() if world == 1 else {"skipped": "measured at N = 1"}
    cpu["reference_python"] = ref
    if not args.no_extras and world == 1:          # CPU arm of the MCTS blocks, bounded samples
        m = arena_port.timed_mcts_sample(seed=args.seed, sims=25, budget_s=min(8.0, args.cpu_seconds))
        b = {"value": m["sims_per_s"], "unit": "sims/s", "cores": m["cores"], "kind": "port",
             "sample": "oracle port of MCTS.py self-play, uniform evaluator, numMCTSSims=25, %d sims in %.1f s on %d processes" % (
                 m["sims"], m["seconds"], m["cores"])}
        if "mcts_uniform_sims_per_s" in ref:
            b["reference_python"] = {"value": ref["mcts_uniform_sims_per_s"], "unit": "sims/s", "cores": 1, "sample": ref["mcts_uniform_sample"]}
        if "mcts_uniform" in extras:
            extras["mcts_uniform"]["cpu_baseline"] = b
        m = arena_port.timed_mcts_nn_sample(seed=args.seed, sims=100, budget_s=min(6.0, args.cpu_seconds))
        b = {"value": m["sims_per_s"], "unit": "sims/s", "cores": m["cores"], "kind": "port",
             "sample": "oracle port of MCTS.py + a float32 CPU forward of the same network per leaf (batch 1, as NNetWrapper.predict), "
                       "numMCTSSims=100, %d sims in %.1f s on %d processes" % (m["sims"], m["seconds"], m["cores"])}
        if "mcts_nn_sims_per_s" in ref:
            b["reference_python"] = {"value": ref["mcts_nn_sims_per_s"], "unit": "sims/s", "cores": 1, "sample": ref["mcts_nn_sample"]}
        for k in ("mcts_nn", "mcts_nn_1m"):
            if k in extras:
                extras[k]["cpu_baseline"] = b
    try:                                  # context only: the same workload in plain C (OpenMP), not the reference's cost
        from oracle import c_oracle
        t0 = time.perf_counter()
        c_steps = c_oracle.timed_steps(400000, PLIES_PER_GAME, args.seed)
        cpu["c_oracle_steps_per_s"] = c_steps / (time.perf_counter() - t0)
    except Exception as exc:              # noqa: BLE001 -- the C oracle is optional test infrastructure
        cpu["c_oracle_steps_per_s"] = None
        cpu["c_oracle_error"] = str(exc)[:100]
    return cpu


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


def run_extras(args, torch, env, dev, n, peaks, prof, rank=0, world=1, dist=None):
    """Secondary kernels of the path timed alone (CUDA events, 3 warm-ups) and the MCTS blocks."""
    out = {}
    peak = peaks["hbm"]

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
    del table

    from nypc_yacht_auction_b200 import mcts_bench as mb
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator, UniformEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet

    def hbm_roofline(sims_per_sec, algo_bytes, kernel, extra=None):
        per_gpu = sims_per_sec / world
        d = {"bound": "hbm", "kernel": kernel, "achieved": per_gpu * algo_bytes / 1e9, "peak": peak, "unit": "GB/s",
             "frac": per_gpu * algo_bytes / 1e9 / peak, "algorithmic_bytes_per_sim": algo_bytes,
             "traffic": prof.get(kernel, {}).get("dram_bytes_per_sim"), "peak_source": peaks["source"],
             "note": "whole-block rate x SURVEY 8(d) bytes per simulation; the tree kernels chase dependent pointers through an "
                     "L2-resident working set, so they sit far below the HBM roof by nature (latency-bound)"}
        d.update(extra or {})
        return d

    # ---- configs[2]: MCTS self-play, numMCTSSims=25, uniform prior (no NN), 4,096 concurrent games per GPU, full episodes
    r = mb.selfplay_block(torch, dev, dist, rank, world, 4096, 25, UniformEvaluator(), args.seed + 1, steps=8, warm=2)
    g = mb.selfplay_block(torch, dev, dist, rank, world, 4096, 25, UniformEvaluator(), args.seed + 1, steps=4, warm=1,
                          fuse_uniform=False, e2e_steps=0)
    r.update({"workload": "configs[2]: MCTS self-play, numMCTSSims=25, uniform prior, 4096 games per GPU, cpuct 1.5, tempThreshold 15; "
                          "one step = one full 48-ply episode of every game with example recording",
              "metric": "mcts_sims_per_sec", "unit": "sims/s", "dtype": "f32", "scaling": "weak",
              "search": "ya_k_mcts_search_uniform: all 25 simulations of a move in one launch, constant-prior nodes",
              "general_path": {"sims_per_sec": g["sims_per_sec"], "ms_per_step": g["ms_per_step"], "gpu_launches": g["gpu_launches"],
                               "search": "ya_mcts_select + ya_mcts_expand<uniform> per simulation (float32 prior rows, the path every "
                                         "non-uniform evaluator takes), replayed as a CUDA graph"},
              "roofline": hbm_roofline(r["sims_per_sec"], mb.ALGO_BYTES_UNIFORM, "ya_k_mcts_search_uniform")})
    out["mcts_uniform"] = r

    # ---- configs[3]: AlphaZero self-play, random-init yacht NNet (H=256, 6 blocks), numMCTSSims=100, 16,384 games per GPU
    torch.manual_seed(0)
    net = YachtPolicyValueNet().to(dev)
    ev = FusedYachtEvaluator(net, 16384, precision="fp16")
    r = mb.selfplay_block(torch, dev, dist, rank, world, 16384, 100, ev, args.seed + 2, steps=2, warm=1)
    wb = mb.wave_breakdown(torch, dev, 16384, 100, ev, args.seed + 2)                       # ply 6: ten-dice score ply, ~3,000 legal moves per leaf
    wb_bid = mb.wave_breakdown(torch, dev, 16384, 100, ev, args.seed + 2, plies_before=8)   # ply 8: first bid of round 3 (202-move rows)
    wb148 = mb.wave_breakdown(torch, dev, 148 * 128, 100, FusedYachtEvaluator(net, 148 * 128, precision="fp16"), args.seed + 2)
    fwd_us, fwd_us_148 = wb["forward_us"], wb148["forward_us"]
    n2 = 2 * 148 * 128                                     # the configs[4] wave: two tiles per CTA on every SM
    wb2 = mb.wave_breakdown(torch, dev, n2, 100, FusedYachtEvaluator(net, n2, precision="fp16"), args.seed + 2)
    wb2_one = mb.wave_breakdown(torch, dev, n2, 100, FusedYachtEvaluator(net, n2, precision="fp16", tiles_per_cta=1), args.seed + 2)
    dense_us = mb.forward_alone(torch, dev, ev, 16384)
    ev.logits = None                                       # drop the dense matrix again (106 MB)
    real_flops = 2.0 * YachtPolicyValueNet.num_macs()
    fwd = {"bound": "tensor", "kernel": "ya_k_forward", "achieved": real_flops * 16384 / fwd_us * 1e-6, "peak": peaks["tensor"],
           "unit": "TFLOP/s", "frac": real_flops * 16384 / fwd_us * 1e-6 / peaks["tensor"], "leaves": 16384, "us_per_launch": fwd_us,
           "flops_per_leaf": real_flops, "peak_source": peaks["source"], "operands": "fp16 (tcgen05 kind::f16), float32 accumulation",
           "how": "CUDA events around the launch inside real simulation waves (ply 6, warm trees): legal logits scattered into the tree "
                  "rows by the policy-head epilogue, no dense logit matrix",
           "dense_logits_variant_us": dense_us,
           "full_machine": {"leaves": 148 * 128, "us_per_launch": fwd_us_148, "achieved": real_flops * 148 * 128 / fwd_us_148 * 1e-6,
                            "frac": real_flops * 148 * 128 / fwd_us_148 * 1e-6 / peaks["tensor"],
                            "note": "one CTA of 128 leaves per SM: 16,384 leaves fill 128 of 148 SMs, 18,944 fill all"},
           "two_tiles_per_cta": {"leaves": n2, "us_per_launch": wb2["forward_us"], "achieved": real_flops * n2 / wb2["forward_us"] * 1e-6,
                                 "frac": real_flops * n2 / wb2["forward_us"] * 1e-6 / peaks["tensor"],
                                 "one_tile_schedule_us": wb2_one["forward_us"], "wave": wb2,
                                 "note": "ya_k_forward2 (waves above one tile per SM, e.g. the configs[4] waves of 37,888 games): each CTA "
                                         "owns two 128-leaf tiles, the tensor core runs one tile's MMAs under the other tile's epilogue; "
                                         "bit-identical to the one-tile schedule (two launch rounds), whose time is given beside it"}}
    bf = mb.selfplay_block(torch, dev, dist, rank, world, 16384, 100, FusedYachtEvaluator(net, 16384, precision="bf16"), args.seed + 2,
                           steps=1, warm=1, e2e_steps=0)
    r.update({"workload": "configs[3]: AlphaZero self-play, random-init YachtNNet (hidden 256, 6 blocks), numMCTSSims=100, 16384 games per GPU; "
                          "one step = one full 48-ply episode of every game with example recording; per simulation wave 3 launches "
                          "(select, whole-network tcgen05 forward, expand with fused softmax + mask), replayed as one CUDA graph",
              "metric": "mcts_sims_per_sec", "unit": "sims/s", "dtype": "fp16 operands / f32 accumulate (the reference's CUDA autocast precision)",
              "scaling": "weak", "nn_flops_per_leaf": real_flops, "wave_breakdown": wb, "wave_breakdown_bid_ply": wb_bid,
              "nn_share_of_time": 0.5 * (wb["nn_share_of_wave"] + wb_bid["nn_share_of_wave"]),
              "bf16_operands": {"sims_per_sec": bf["sims_per_sec"], "ms_per_step": bf["ms_per_step"]},
              "roofline": hbm_roofline(r["sims_per_sec"], mb.ALGO_BYTES_NN_F32, "ya_mcts_wave",
                                       {"algorithmic_bytes_per_sim_16bit": mb.ALGO_BYTES_NN_16,
                                        "frac_16bit": r["sims_per_sec"] / world * mb.ALGO_BYTES_NN_16 / 1e9 / peak}),
              "roofline_forward": fwd})
    out["mcts_nn"] = r
    out["nn_forward"] = fwd

    # ---- configs[4]: 1,048,576 games (all ranks together), numMCTSSims=100, waves of 16,384 games per tree pool: STRONG scaling
    if not args.no_1m:
        total = args.games_1m
        r = mb.selfplay_1m(torch, dev, dist, rank, world, lambda m: FusedYachtEvaluator(net, m, precision="fp16"), total, 100,
                           args.wave_1m, args.seed + 3)
        r.update({"workload": "configs[4]: %d concurrent games sharded over %d GPU(s) by global game id, numMCTSSims=100, random-init "
                              "YachtNNet, waves of %d games on one tree pool per GPU" % (total, world, r["wave_games"]),
                  "metric": "mcts_sims_per_sec", "unit": "sims/s", "dtype": "fp16 operands / f32 accumulate",
                  "roofline": hbm_roofline(r["sims_per_sec"], mb.ALGO_BYTES_NN_F32, "ya_mcts_wave",
                                           {"frac_16bit": r["sims_per_sec"] / world * mb.ALGO_BYTES_NN_16 / 1e9 / peak}),
                  "e2e": {"value": r["sims_per_sec"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": r["d2h_bytes_per_gpu"],
                          "note": "the block itself is end to end: games are dealt on device from the seed (no input but the seed), every "
                                  "wave's examples are copied to pinned host memory inside the timed region"}})
        out["mcts_nn_1m"] = r

    # ---- drop-in MCTS class (one game, host evaluator called per leaf exactly like MCTS.py:86)
    if rank == 0:
        out["dropin_mcts"] = dropin_mcts_rate(torch, dev)
    if world > 1:
        out["shard_invariance"] = mb.shard_invariance_check(torch, dev, dist, rank, world, net)
    return out


def dropin_mcts_rate(torch, dev, sims=25, plies=12):
    """The drop-in `MCTS(game, nnet, args).getActionProb` for ONE game with a host-side uniform evaluator: what a caller of
    the reference's API gets without batching (kernel launch + host sync per simulation)."""
    import numpy as np
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.mcts import MCTS

    class UniformNet:
        pi = np.full(ACTION_SIZE, np.float32(1.0) / np.float32(ACTION_SIZE), dtype=np.float32)

        def predict(self, board):
            return self.pi, np.float32(0.0)

    class Args(dict):
        __getattr__ = dict.__getitem__

    g = YachtGame(seed=0)
    mcts = MCTS(g, UniformNet(), Args(numMCTSSims=sims, cpuct=1.5))
    board, cur = g.getInitBoard(), 1
    rng = np.random.RandomState(0)
    t0, done = None, 0
    for ply in range(plies + 2):
        if ply == 2:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        pi = mcts.getActionProb(g.getCanonicalForm(board, cur), temp=1)
        if ply >= 2:
            done += sims
        board, cur = g.getNextState(board, cur, int(rng.choice(len(pi), p=pi)))
    dt = time.perf_counter() - t0
    return {"sims_per_sec": done / dt, "games": 1, "num_mcts_sims": sims, "plies_timed": plies,
            "note": "drop-in MCTS class, one game, host evaluator per leaf (MCTS.py:86): latency-bound by one launch + one sync per "
                    "simulation; the batched searcher above is the throughput path"}


def main():
    args = parse_args()
    wd = float(os.environ.get("YA_BENCH_WATCHDOG", "0") or 0)   # seconds; > 0: dump every thread's stack to stderr and exit if still running
    if wd > 0:
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
