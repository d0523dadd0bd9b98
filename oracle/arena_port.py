"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU baseline driver.

Restates the reference's config-1 workload, ``Arena.playGame`` with ``RandomYachtPlayer`` on both
sides (Arena.py:30-93, yacht/YachtPlayers.py:174-183), on top of oracle/yacht_rules.py, keeping
the reference's call structure (two legal-mask evaluations per ply: one by the player, one by
the arena's legality assert) so its cost profile is that of the reference's Python path.  Dice and
picks come from the Philox protocol, so a game is the same game the GPU plays.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import philox
from . import yacht_rules as yr


def play_game(seed, game_id, episode=0):
    """One full game; returns (plies, outcome for player 1)."""
    board = yr.new_game(philox.Draw(seed, game_id, episode, 0, philox.TAG_INIT))
    cur, ply = 1, 0
    while yr.outcome(board, cur) == 0:                               # Arena.py:49
        canon = yr.canonical(board, cur)                             # Arena.py:55-56
        legal = np.nonzero(yr.legal_mask(canon, 1))[0]               # YachtPlayers.py:181-182
        pick = philox.Draw(seed, game_id, episode, ply, philox.TAG_ACTION).pick(len(legal)) if len(legal) else 0
        action = int(legal[pick]) if len(legal) else 0
        valids = yr.legal_mask(yr.canonical(board, cur), 1)          # Arena.py:58-64
        assert valids[action] > 0
        board, cur = yr.next_state(board, cur, action, philox.Draw(seed, game_id, episode, ply, philox.TAG_REAL))
        ply += 1
    return ply, yr.outcome(board, 1)


def _worker(args):
    seed, first, count, budget_s = args
    t0 = time.perf_counter()
    steps = games = 0
    for g in range(first, first + count):
        p, _ = play_game(seed, g)
        steps += p
        games += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return steps, games, time.perf_counter() - t0


def timed_sample(seed=0, budget_s=10.0, procs=None, games_per_proc=100000):
    """Plays games on `procs` processes (default: all host cores) for about budget_s seconds.
    Returns dict(steps_per_s, games, steps, cores, seconds)."""
    procs = procs or os.cpu_count() or 1
    jobs = [(seed, i * games_per_proc, games_per_proc, budget_s) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_worker(jobs[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    games = sum(r[1] for r in res)
    # aggregate throughput: every worker ran for its own measured time
    rate = sum(r[0] / r[2] for r in res)
    return {"steps_per_s": rate, "games": games, "steps": steps, "cores": procs, "seconds": wall}


# ---------------------------------------------------------------------------- MCTS baseline
def _mcts_worker(args):
    from . import mcts_oracle
    seed, first, sims, budget_s = args
    t0 = time.perf_counter()
    done = plies = 0
    g = first
    while time.perf_counter() - t0 < budget_s:
        trace, *_ = mcts_oracle.self_play_game(mcts_oracle.uniform_evaluator, sims, 1.5, seed, g, max_plies=12)
        plies += len(trace)
        done += len(trace) * sims
        g += 1
    return done, plies, time.perf_counter() - t0


def timed_mcts_sample(seed=0, sims=25, budget_s=8.0, procs=None):
    """Oracle restatement of MCTS.py self-play (uniform evaluator, numMCTSSims=sims, first 12 plies of each
    game) on `procs` processes for about budget_s seconds.  Returns dict(sims_per_s, cores, ...)."""
    procs = procs or os.cpu_count() or 1
    jobs = [(seed, 1000 * i, sims, budget_s) for i in range(procs)]
    if procs == 1:
        res = [_mcts_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_mcts_worker, jobs)
    return {"sims_per_s": sum(r[0] / r[2] for r in res), "steps_per_s": sum(r[1] / r[2] for r in res), "cores": procs,
            "sims": sum(r[0] for r in res), "seconds": max(r[2] for r in res)}


# ---------------------------------------------------------------------------- MCTS + network baseline
def _mcts_nn_worker(args):
    """MCTS.py self-play with NNetWrapper.predict on the CPU (yacht/NNet.py:177-195: state_to_vec, one float32
    forward of YachtNNet for ONE board, softmax) -- the reference's configs[3] cost profile, batch 1."""
    import torch
    from . import mcts_oracle
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet       # same architecture as yacht/pytorch/YachtNNet.py
    seed, first, sims, budget_s = args
    torch.set_num_threads(1)
    torch.manual_seed(0)
    net = YachtPolicyValueNet().eval()

    def predict(board):
        x = torch.from_numpy(yr.features(board)).unsqueeze(0)
        with torch.no_grad():
            logits, v = net(x)
        return torch.softmax(logits, dim=1)[0].numpy(), np.float32(v.reshape(-1)[0].item())

    t0 = time.perf_counter()
    done = 0
    g = first
    while time.perf_counter() - t0 < budget_s:
        trace, *_ = mcts_oracle.self_play_game(predict, sims, 1.5, seed, g, max_plies=2)
        done += len(trace) * sims
        g += 1
    return done, time.perf_counter() - t0


def timed_mcts_nn_sample(seed=0, sims=100, budget_s=6.0, procs=None):
    """Oracle MCTS + float32 CPU forward per leaf (batch 1, one thread per process), first 2 plies of each game."""
    procs = procs or os.cpu_count() or 1
    jobs = [(seed, 1000 * i, sims, budget_s) for i in range(procs)]
    if procs == 1:
        res = [_mcts_nn_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_mcts_nn_worker, jobs)
    return {"sims_per_s": sum(r[0] / r[1] for r in res), "cores": procs, "sims": sum(r[0] for r in res),
            "seconds": max(r[1] for r in res)}
