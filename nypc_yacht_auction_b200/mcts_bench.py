"""MCTS self-play blocks of bench.py: BASELINE.json configs[2] (uniform prior, 25 sims, 4,096 games), configs[3]
(AlphaZero self-play, random-init YachtNNet, 100 sims, 16,384 games) and configs[4] (1,048,576 games, 100 sims, sharded
over the ranks and played in waves on one tree pool per GPU -- strong scaling).

A bench "step" of these blocks is ONE FULL 48-ply episode of every game (Coach.executeEpisode, Coach.py:34-72): per ply
numMCTSSims simulations, example recording, move sampling, transition.  Device time = CUDA events on the launch stream,
barrier + synchronize on both sides, max over ranks.  The e2e figures go through the host-buffer call
(BatchedSelfPlay.execute_episodes_host: start boards in pinned host memory -> examples in pinned host memory).
"""
from __future__ import annotations

import os
import sys
import time

PLIES = 48
# SURVEY.md section 8(d), algorithmic bytes per simulation: ~7.5 KB (uniform prior: prior rows + edges on a depth-2 path),
# ~33 KB with the network (+ features + float32 logits written and read once).  The "engine" figure is what THIS engine moves
# by design: 236 B features + the L legal logits written once by the forward's epilogue and read once by expand as 16-bit
# values (L ~ 910 on average) + L / 4 B of group maxima and visited bits + ~3.6 KB of node / edge traffic on the path + value.
ALGO_BYTES_UNIFORM = 7500
ALGO_BYTES_NN_F32 = 33000
ALGO_BYTES_NN_16 = 236 + 2 * 910 + 2 * 910 + 910 // 4 + 3600 + 200


def _fence(torch, dev, dist, world):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)


def _max_ranks(torch, dev, dist, world, x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _launches_per_episode(per_sim, sims):
    # per move: features (+ canonical form), root counts, sparse root policy, overflow max, pick, transition, ply bump ~ 8
    # small launches, plus the search: ONE launch (uniform prior, whole search fused: per_sim = 0), 2 per simulation
    # (select, expand: uniform prior on the general path) or 3 (select, forward, expand: network)
    return PLIES * (8 + (per_sim * sims if per_sim else 1))


def selfplay_block(torch, dev, dist, rank, world, n, sims, evaluator, seed, steps, warm, use_graph=True, fuse_uniform=True,
                   e2e_steps=1, arena_mb=None):
    """`steps` full episodes of n games per rank after `warm` warm-up episodes."""
    from .coach import BatchedSelfPlay
    sp = BatchedSelfPlay(n, sims, cpuct=1.5, evaluator=evaluator, temp_threshold=15, seed=seed, game_base=rank * n, device=dev,
                         arena_mb_per_game=arena_mb, record_examples=True)
    uniform = getattr(sp.mcts.evaluator, "uniform", False)
    sp.mcts.fuse_uniform = bool(fuse_uniform)
    if use_graph and not (uniform and fuse_uniform):
        sp.mcts.capture_graph()

    peak_nodes = [0]

    def episode():
        sp.execute_episodes()
        peak_nodes[0] = max(peak_nodes[0], int(sp.mcts.pool.node_counts().max().item()))   # live nodes after the last ply (pruned per round)
        sp.next_episode()

    for _ in range(warm):
        episode()
    _fence(torch, dev, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        episode()
    e1.record()
    _fence(torch, dev, dist, world)
    ms = _max_ranks(torch, dev, dist, world, e0.elapsed_time(e1))
    nodes = peak_nodes[0]
    out = {"ms_per_step": ms / steps, "steps": steps, "warmup": warm,
           "sims_per_sec": world * n * sims * PLIES * steps / (ms * 1e-3),
           "game_steps_per_sec": world * n * PLIES * steps / (ms * 1e-3),
           "gpu_launches": steps * _launches_per_episode(0 if (uniform and fuse_uniform) else (2 if uniform else 3), sims),
           "pool_gb": sp.mcts.pool.bytes() / 1e9, "nodes_live_at_episode_end": nodes}
    if e2e_steps:
        host = sp.host_buffers()
        host["boards"].copy_(sp.env.states)                         # getInitBoard positions of the next episode, kept on the host
        host["players"].copy_(sp.env.players)
        h2d, d2h = sp.execute_episodes_host(host)                   # warm (pinned buffers touched)
        _fence(torch, dev, dist, world)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h2d, d2h = sp.execute_episodes_host(host)
        dt = _max_ranks(torch, dev, dist, world, time.perf_counter() - t0)
        out["e2e"] = {"value": world * n * sims * PLIES * e2e_steps / dt, "unit": "sims/s",
                      "game_steps_per_sec": world * n * PLIES * e2e_steps / dt,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                      "api": "BatchedSelfPlay.execute_episodes_host: start boards + movers from pinned host memory -> 48 plies of "
                             "MCTS self-play -> feature rows, sparse root visit counts, labels, outcomes into pinned host memory",
                      "timing": "host wall clock around synchronous calls, max over ranks"}
        del host
    sp.mcts.check_errors()
    del sp
    torch.cuda.empty_cache()
    return out


def wave_breakdown(torch, dev, n, sims, evaluator, seed, plies_before=6, waves=60):
    """Device time of the three kernels of one network simulation wave (select, forward, expand), each bracketed by CUDA
    events on the launch stream, at a mid-game ply with warm trees; the NN share of the wave follows."""
    from . import _lib
    from .engine import BatchedYacht
    from .mcts import BatchedMCTS
    env = BatchedYacht(n, seed=seed, game_base=0, device=dev)
    m = BatchedMCTS(env, sims, 1.5, evaluator=evaluator)
    for _ in range(plies_before):
        m.play_ply()
    grp, lib, s = m.groups[0], m.lib, _lib.current_stream()
    ev = grp.evaluator
    tot = [0.0, 0.0, 0.0]
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(waves)]
    for sim in range(waves):
        a = marks[sim]
        a[0].record()
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                      env.seed, env.game_base, sim, None, None, m.cpuct, None, _lib.ptr(grp.features),
                                      _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                      _lib.ptr(m.err_flag), s), "ya_mcts_select")
        a[1].record()
        pi, v = ev(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
        a[2].record()
        _lib.check(lib.ya_mcts_expand_logits(grp.ref, None, 1 if m.rows == 2 else 0, 0, _lib.ptr(ev.last_row_max), _lib.ptr(v), None,
                                             _lib.ptr(m.err_flag), s), "ya_mcts_expand_logits")
        a[3].record()
    torch.cuda.synchronize(dev)
    m.check_errors()
    for a in marks[5:]:
        for i in range(3):
            tot[i] += a[i].elapsed_time(a[i + 1])
    k = len(marks) - 5
    sel, fwd, exp = (1e3 * t / k for t in tot)
    del m, env
    torch.cuda.empty_cache()
    return {"select_us": sel, "forward_us": fwd, "expand_us": exp, "nn_share_of_wave": fwd / (sel + fwd + exp),
            "how": "CUDA events around each launch, %d waves at ply %d, eager (the graphed wave has no gaps between them)" % (k, plies_before)}


def forward_alone(torch, dev, evaluator, leaves, reps=50):
    x = torch.rand((leaves, 59), device=dev)
    for _ in range(5):
        evaluator(x)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(reps):
        evaluator(x)
    f1.record()
    torch.cuda.synchronize(dev)
    return f0.elapsed_time(f1) * 1000.0 / reps


def wave_plan(games, wave_cap, tile=128):
    """(number of waves, games per wave) for a shard of `games` games on a tree pool of `wave_cap` games: as few waves as the
    pool allows, all of the same size, padded to whole forward tiles."""
    n_waves = max(1, (games + wave_cap - 1) // wave_cap)
    wave = ((games + n_waves - 1) // n_waves + tile - 1) // tile * tile
    return n_waves, wave


def selfplay_1m(torch, dev, dist, rank, world, evaluator_factory, total, sims, wave_cap, seed):
    """configs[4]: `total` games split over the ranks by global game id (dist.shard_range), every rank plays its shard as
    consecutive waves on ONE tree pool (coach.self_play_in_waves; simulation wave replayed as a CUDA graph) and copies every
    wave's examples to pinned host memory.  STRONG scaling: total work is fixed.  The pool, the graph and the pinned buffers
    are built before the timed region (one-time setup, like loading the weights); every wave's pool reset, re-deal, search,
    example recording and device->host copy are inside it."""
    from .coach import make_wave_player, self_play_in_waves
    from .dist import shard_range
    first, last = shard_range(total, rank, world)
    mine = last - first
    # Wave size.  A forward launch costs about the same for anything up to two 128-leaf tiles per SM (2 x 148 x 128 = 37,888
    # leaves = `wave_cap`: the two-tiles-per-CTA kernel), and the tree kernels scale with the games, so: as few waves as the
    # pool allows, all of the same size (one pool, one captured graph), padded to whole 128-leaf tiles.
    wave = wave_plan(mine, wave_cap)[1]
    sp = make_wave_player(mine, wave, sims, evaluator_factory(wave), first_game=first, use_graph=True, seed=seed, device=dev)
    names = ("features", "actions", "counts", "value", "result_p1")
    host = {"features": torch.empty((sp.PLIES, sp.n, 59), dtype=torch.float32, pin_memory=True),
            "actions": torch.empty((sp.PLIES, sp.n, sp.k), dtype=torch.int16, pin_memory=True),
            "counts": torch.empty((sp.PLIES, sp.n, sp.k), dtype=torch.int32, pin_memory=True),
            "value": torch.empty((sp.PLIES, sp.n), dtype=torch.float32, pin_memory=True),
            "result_p1": torch.empty(sp.n, dtype=torch.float32, pin_memory=True)}
    tally = {"d2h": 0, "waves": 0}

    def on_wave(w, ex):
        for name in names:
            gd = 0 if ex[name].dim() == 1 else 1                 # a ragged last wave fills only its share of the buffer
            dst = host[name].narrow(gd, 0, ex[name].shape[gd])
            dst.copy_(ex[name], non_blocking=True)
            tally["d2h"] += dst.numel() * dst.element_size()
        tally["waves"] += 1
        if os.environ.get("YA_BENCH_PROGRESS"):
            print("[selfplay_1m] rank %d wave %d done after %.1f s" % (rank, w, time.perf_counter() - t0), file=sys.stderr, flush=True)

    _fence(torch, dev, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    p1, p2, dr = self_play_in_waves(mine, wave, sims, None, first_game=first, on_wave=on_wave, sp=sp)
    e1.record()
    _fence(torch, dev, dist, world)
    wall = _max_ranks(torch, dev, dist, world, time.perf_counter() - t0)
    ms = _max_ranks(torch, dev, dist, world, e0.elapsed_time(e1))
    res = torch.tensor([p1, p2, dr], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(res)
    res = [int(x) for x in res.tolist()]
    assert sum(res) == total, "every game must finish"
    out = {"games": total, "games_per_gpu": mine, "wave_games": wave, "waves_per_gpu": tally["waves"],
           "last_wave_live_games": mine - (tally["waves"] - 1) * wave, "num_mcts_sims": sims,
           "seconds": ms * 1e-3, "wall_seconds": wall, "sims_per_sec": total * PLIES * sims / (ms * 1e-3),
           "game_steps_per_sec": total * PLIES / (ms * 1e-3), "scaling": "strong",
           "outcomes_p1_p2_draw": res, "d2h_bytes_per_gpu": tally["d2h"], "pool_gb": sp.mcts.pool.bytes() / 1e9,
           "gpu_launches_per_gpu": tally["waves"] * _launches_per_episode(3, sims),
           "wave_choice": "as few equal waves as the tree pool (capacity %d games = two 128-leaf tiles per SM, one launch of the two-tiles-per-CTA "
                          "forward) allows, padded to whole tiles; the last wave's surplus slots play games that are dropped" % wave_cap,
           "includes": "per wave: pool reset + re-deal, 48 plies x numMCTSSims simulations, example recording, device->host copy of the "
                       "examples into pinned memory; excludes the one-time setup (pool allocation, CUDA-graph capture, pinned buffers)"}
    del host, sp
    torch.cuda.empty_cache()
    return out


def shard_invariance_check(torch, dev, dist, rank, world, net, total=64, sims=16, seed=77):
    """N > 1: `total` games sharded over the ranks, each shard played as two graphed waves with the network evaluator,
    all-gathered over NCCL and compared on rank 0 with one single-process batch of all games.  Returns "ok" or raises."""
    from .coach import BatchedSelfPlay, self_play_in_waves
    from .dist import allgather_examples, shard_range
    from .mcts import FusedYachtEvaluator
    first, last = shard_range(total, rank, world)
    per = last - first
    assert per % 2 == 0
    parts = []
    self_play_in_waves(per, per // 2, sims, FusedYachtEvaluator(net, per // 2), first_game=first, use_graph=True, seed=seed, device=dev,
                       on_wave=lambda w, ex: parts.append({k: v.clone() for k, v in ex.items()}))
    names = ("features", "actions", "counts", "value", "result_p1")
    mine = {k: torch.cat([p[k] for p in parts], dim=0 if parts[0][k].dim() == 1 else 1) for k in names}
    full = allgather_examples(mine)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if rank == 0:
        ref = BatchedSelfPlay(total, sims, evaluator=FusedYachtEvaluator(net, total), seed=seed, game_base=0, device=dev).execute_episodes()
        ok[0] = int(all(torch.equal(full[k], ref[k]) for k in names))
    dist.broadcast(ok, src=0)
    if int(ok.item()) != 1:
        raise AssertionError("sharded self-play differs from the single-process run")
    torch.cuda.empty_cache()
    return "ok"
