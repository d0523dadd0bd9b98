#!/usr/bin/env python
"""Times the UNMODIFIED reference (iyioon/NYPC-Yacht-Auction) on the host CPU through its own public API.  Run as a
separate process with the reference's directory as argv[1] (baseline/_ref on the GPU box, /root/reference in the build
container); prints one JSON object.  Nothing of this repo's engine is imported here.

    python baseline/run_reference.py <ref_dir> arena   <games>           # BASELINE.json configs[0], BASELINE.md section 4.2
    python baseline/run_reference.py <ref_dir> mcts    <sims> <games>    # MCTS.py + uniform evaluator (section 4.3)
    python baseline/run_reference.py <ref_dir> mcts_nn <sims> <games>    # MCTS.py + NNetWrapper(cuda=False) (section 4.4)
    python baseline/run_reference.py <ref_dir> arena_for <seconds> <seed>  # as many Arena games as fit (one worker of the all-core arm)
"""
import json
import logging
import os
import sys
import time


def main():
    ref_dir, mode = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ref_dir)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    logging.disable(logging.CRITICAL)                      # Arena logs every game's totals (Arena.py:78-84)
    import numpy as np
    from yacht.YachtGame import YachtGame
    from yacht.YachtPlayers import RandomYachtPlayer
    from Arena import Arena
    import tqdm as _tqdm_mod
    import Arena as arena_mod
    import Coach as coach_mod
    quiet = lambda it, **kw: it                            # tqdm progress bars off (Arena.py:110,121; Coach.py:91)
    arena_mod.tqdm = quiet
    coach_mod.tqdm = quiet

    if mode in ("arena", "arena_for"):
        seed = int(sys.argv[4]) if mode == "arena_for" else 0
        g = YachtGame(seed=seed)
        arena = Arena(RandomYachtPlayer(g).play, RandomYachtPlayer(g).play, g)
        if mode == "arena":
            games = int(sys.argv[3])
            t0 = time.perf_counter()
            one, two, draws = arena.playGames(games)
            dt = time.perf_counter() - t0
        else:
            budget = float(sys.argv[3])
            games = one = two = draws = 0
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < budget:
                a, b, d = arena.playGames(4)
                one, two, draws, games = one + a, two + b, draws + d, games + 4
            dt = time.perf_counter() - t0
        assert one + two + draws == games
        print(json.dumps({"mode": mode, "games": games, "steps": 48 * games, "seconds": dt, "games_per_s": games / dt,
                          "steps_per_s": 48 * games / dt, "p1": one, "p2": two, "draws": draws,
                          "api": "Arena(RandomYachtPlayer(g).play, RandomYachtPlayer(g).play, g).playGames(%d), YachtGame(seed=%d)" % (games, seed)}))
        return

    sims, games = int(sys.argv[3]), int(sys.argv[4])
    from MCTS import MCTS
    from utils import dotdict
    import torch
    args = dotdict({"numMCTSSims": sims, "cpuct": 1.5, "tempThreshold": 15, "lr": 2e-3, "weight_decay": 1e-4, "epochs": 1,
                    "batch_size": 512, "vloss_weight": 1.5, "cuda": False, "hidden": 256, "nblocks": 6, "dropout": 0.3,
                    "numItersForTrainExamplesHistory": 5, "maxlenOfQueue": 200000})
    g = YachtGame(seed=0)
    if mode == "mcts":
        class UniformNet:                                  # P = 1/3226 (float32), v = 0: BASELINE.json configs[2]
            def __init__(self, game=None, args=None):
                self.pi = np.full(3226, np.float32(1.0) / np.float32(3226), dtype=np.float32)

            def predict(self, board):
                return self.pi.copy(), np.float32(0.0)
        net = UniformNet()
    else:
        from yacht.NNet import NNetWrapper
        torch.manual_seed(0)
        torch.set_num_threads(1)
        net = NNetWrapper(g, args)
    coach = coach_mod.Coach(g, net, args)
    plies = 0
    t0 = time.perf_counter()
    for _ in range(games):
        coach.mcts = MCTS(g, net, args)                    # Coach.py:93: a fresh tree per episode
        plies += len(coach.executeEpisode())
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": mode, "games": games, "plies": plies, "sims": plies * sims, "seconds": dt,
                      "sims_per_s": plies * sims / dt, "steps_per_s": plies / dt, "num_mcts_sims": sims,
                      "api": "Coach.executeEpisode with MCTS(game, %s, numMCTSSims=%d, cpuct=1.5)" % (
                          "uniform evaluator" if mode == "mcts" else "NNetWrapper(cuda=False, hidden=256, nblocks=6)", sims)}))


if __name__ == "__main__":
    main()
