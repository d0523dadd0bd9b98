"""Self-play surface of the reference's Coach (/root/reference/Coach.py:34-72).

* ``Coach.executeEpisode`` -- the reference's one-game loop, unchanged in behaviour, for use with the
  drop-in ``YachtGame`` / ``MCTS`` (global numpy RNG for the move sampling, like Coach.py:65).
* ``BatchedSelfPlay`` -- the same episode for thousands of games in lock-step on the GPU: every
  ply = numMCTSSims x (select, ONE batched evaluator call, expand/backup), then device-side move
  sampling and the fused transition.  Training examples stay on the device as compact tensors
  (feature rows, sparse visit counts, outcomes) instead of pickled Python objects.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import ACTION_SIZE, FEATURE_SIZE, BatchedYacht
from .mcts import MCTS, BatchedMCTS


class Coach:
    """Coach.py:17-72 (self-play part).  ``learn`` (training / arena gate, Coach.py:74-139) is
    orchestration outside the accelerated path: use the reference's Coach with this package's
    YachtGame and MCTS for it."""

    def __init__(self, game, nnet, args):
        self.game = game
        self.nnet = nnet
        self.args = args
        self.mcts = MCTS(self.game, self.nnet, self.args)
        self.trainExamplesHistory = []
        self.skipFirstSelfPlay = False

    def executeEpisode(self):
        """Coach.py:34-72."""
        trainExamples = []
        board = self.game.getInitBoard()
        self.curPlayer = 1
        episodeStep = 0
        while True:
            episodeStep += 1
            canonicalBoard = self.game.getCanonicalForm(board, self.curPlayer)
            temp = int(episodeStep < self.args.tempThreshold)
            pi = self.mcts.getActionProb(canonicalBoard, temp=temp)
            sym = self.game.getSymmetries(canonicalBoard, pi)
            for b, p in sym:
                trainExamples.append([b, self.curPlayer, p, None])
            action = np.random.choice(len(pi), p=pi)
            board, self.curPlayer = self.game.getNextState(board, self.curPlayer, action)
            r = self.game.getGameEnded(board, self.curPlayer)
            if r != 0:
                return [(x[0], x[2], r * ((-1) ** (x[1] != self.curPlayer))) for x in trainExamples]


class BatchedSelfPlay:
    """n concurrent episodes; game g of rank r is global game ``game_base + g``."""

    PLIES = 48

    def __init__(self, n, num_sims, cpuct=1.5, evaluator=None, temp_threshold=15, seed=0, game_base=0,
                 device="cuda", arena_mb_per_game=None, record_examples=True, max_edges=None, record_states=False):
        self.env = BatchedYacht(n, seed=seed, game_base=game_base, device=device)
        self.mcts = BatchedMCTS(self.env, num_sims, cpuct, evaluator, temp_threshold, arena_mb_per_game)
        self.n = n
        self.record = record_examples
        self.record_states = bool(record_states) and self.record      # packed canonical boards: examples.to_reference_examples
        self.k = int(max_edges or min(ACTION_SIZE, 2 * num_sims))
        d = self.env.device
        if self.record:
            self.ex_features = torch.zeros((self.PLIES, n, FEATURE_SIZE), dtype=torch.float32, device=d)
            self.ex_actions = torch.zeros((self.PLIES, n, self.k), dtype=torch.int16, device=d)
            self.ex_counts = torch.zeros((self.PLIES, n, self.k), dtype=torch.int32, device=d)
            self.ex_players = torch.zeros((self.PLIES, n), dtype=torch.int8, device=d)
            self.ex_overflow = torch.zeros(n, dtype=torch.int32, device=d)       # per ply (rewritten by every root_sparse)
            self.ex_overflow_max = torch.zeros(n, dtype=torch.int32, device=d)   # largest edge count that did not fit k, per game
            if self.record_states:
                self.ex_states = torch.zeros((self.PLIES, 2, n, 4), dtype=torch.int32, device=d)

    def play_ply(self, t):
        env, m = self.env, self.mcts
        if self.record:
            if self.record_states:
                env.features(canonical_states=env.canonical(out=self.ex_states[t]), out=self.ex_features[t])
            else:
                env.features(out=self.ex_features[t])                # state_to_vec of the canonical root
            self.ex_players[t].copy_(env.players)
        m.search()
        m.root_counts()
        if self.record:                                              # the visited root edges, sorted by action
            m.root_sparse(self.ex_actions[t], self.ex_counts[t], self.ex_overflow)
            torch.maximum(self.ex_overflow_max, self.ex_overflow, out=self.ex_overflow_max)
        actions = m.pick_actions()
        env.next_state(actions, check=False)
        return actions

    def execute_episodes(self):
        """One full episode for every game.  Returns dict(features[T,n,59], actions[T,n,k], counts[T,n,k],
        value[T,n]) with value = +-1 / 1e-4 from the mover's view (Coach.py:69-72)."""
        if self.record:
            self.ex_overflow_max.zero_()
        for t in range(self.PLIES):
            self.play_ply(t)
        self.mcts.check_errors()
        if self.record:
            # a root keeps the edges earlier searches of the same round visited, so it can hold more than numMCTSSims of
            # them; a pi target cut to the k lowest actions would be silently wrong -> refuse it
            worst = int(self.ex_overflow_max.max().item())
            if worst:
                raise ValueError("a root had %d visited edges but the example buffers hold max_edges=%d per ply: "
                                 "construct BatchedSelfPlay with a larger max_edges" % (worst, self.k))
        env = self.env
        ones = torch.ones_like(env.players)
        result_p1 = env.game_ended(players=ones)                     # from player 1's view
        assert bool((result_p1 != 0).all()), "every game ends after 48 plies"
        out = None
        if self.record:
            # r * (-1)**(player != curPlayer) with curPlayer == 1 at the end of round 13 (Coach.py:69-72):
            # +-1 for a decided game and +-1e-4 for a draw, both = result_p1 * player
            value = result_p1.unsqueeze(0) * self.ex_players.float()
            out = {"features": self.ex_features, "actions": self.ex_actions, "counts": self.ex_counts, "value": value,
                   "result_p1": result_p1}
            if self.record_states:
                out["states"] = self.ex_states
        return out

    def next_episode(self):
        self.mcts.new_episode()

    # ------------------------------------------------------------------ host-buffer form (the end-to-end call)
    def host_buffers(self):
        """Pinned HOST buffers for execute_episodes_host: the start boards (two uint4 planes, int32[2, n, 4]) + movers in,
        the training examples of Coach.executeEpisode (Coach.py:66-72) out."""
        n, k, pin = self.n, self.k, dict(pin_memory=True)
        return {
            "boards": torch.zeros((2, n, 4), dtype=torch.int32, **pin), "players": torch.ones(n, dtype=torch.int8, **pin),
            "features": torch.zeros((self.PLIES, n, FEATURE_SIZE), dtype=torch.float32, **pin),
            "actions": torch.zeros((self.PLIES, n, k), dtype=torch.int16, **pin),
            "counts": torch.zeros((self.PLIES, n, k), dtype=torch.int32, **pin),
            "value": torch.zeros((self.PLIES, n), dtype=torch.float32, **pin),
            "result_p1": torch.zeros(n, dtype=torch.float32, **pin),
        }

    def execute_episodes_host(self, host, new_trees=True):
        """One self-play episode for every game with HOST-resident inputs and outputs: the start boards in
        host["boards"] / host["players"] (e.g. getInitBoard positions the caller keeps) go host -> device, the 48
        plies run, and the examples (feature rows, sparse root visit counts, outcome labels) come back into the
        pinned buffers of `host`.  Returns (h2d_bytes, d2h_bytes).  Synchronises."""
        env = self.env
        if new_trees:
            self.mcts.pool.reset()
        env.states.copy_(host["boards"], non_blocking=True)
        env.players.copy_(host["players"], non_blocking=True)
        env.ply.zero_()
        ex = self.execute_episodes()
        names = ("features", "actions", "counts", "value", "result_p1")
        for name in names:
            host[name].copy_(ex[name], non_blocking=True)
        torch.cuda.synchronize(env.device)
        h2d = host["boards"].numel() * 4 + host["players"].numel()
        d2h = sum(host[name].numel() * host[name].element_size() for name in names)
        return h2d, d2h

    @staticmethod
    def dense_policy(actions, counts):
        """Sparse (actions, counts) rows -> float64 pi[.., 3226] = counts / sum (MCTS.py:51-54, temp 1)."""
        shape = actions.shape[:-1]
        pi = torch.zeros(shape + (ACTION_SIZE,), dtype=torch.float64, device=actions.device)
        pi.scatter_add_(-1, actions.long() & 0xFFFF, counts.double())
        return pi / pi.sum(-1, keepdim=True)


class BatchedArena:
    """The arena gate of Coach.learn (Coach.py:118-126, Arena.playGames Arena.py:95-130) for two MCTS
    populations: `n` games with evaluator A as player 1 and `n` games with the sides swapped, every player
    moving greedily on its own visit counts (temp = 0: np.argmax(getActionProb(x, temp=0)), ties broken
    uniformly).  Each side keeps its own trees, like the reference's separate pmcts / nmcts objects.
    All games follow the same ply schedule, so at every ply one population searches a whole batch.
    A seat can also be a scripted device player -- pass "greedy" (GreedyYachtPlayer, yacht/YachtPlayers.py:186-214)
    or "random" (RandomYachtPlayer, :174-183) instead of an evaluator -- which is how a network is measured
    against the reference's heuristic opponent (Arena(mcts_player, GreedyYachtPlayer(g).play, g))."""

    PLIES = 48

    def __init__(self, n, num_sims, evaluator_a, evaluator_b, cpuct=1.5, seed=0, game_base=0, device="cuda",
                 arena_mb_per_game=None):
        self.n = n
        self.envs = [BatchedYacht(n, seed=seed, game_base=game_base, device=device),
                     BatchedYacht(n, seed=seed, game_base=game_base + n, device=device)]
        evs = (evaluator_a, evaluator_b)
        # searchers[e][k]: the searcher of evaluator k on env e; env 0: A is player +1, env 1: B is player +1
        self.searchers = [[evs[k] if isinstance(evs[k], str) else
                           BatchedMCTS(env, num_sims, cpuct, evs[k], temp_threshold=0, arena_mb_per_game=arena_mb_per_game)
                           for k in range(2)] for env in self.envs]
        for ev in evs:
            if isinstance(ev, str) and ev not in ("greedy", "random"):
                raise ValueError("scripted seat must be 'greedy' or 'random', got %r" % (ev,))

    def play_games(self):
        """Returns (a_wins, b_wins, draws) over the 2n games."""
        for t in range(self.PLIES):
            for e, env in enumerate(self.envs):
                mover = int(env.players[0].item())                  # identical for every game of the batch
                k = (0 if mover == 1 else 1) ^ e                    # which evaluator owns this seat
                m = self.searchers[e][k]
                if isinstance(m, str):                              # scripted seat
                    env.next_state(env.greedy_actions() if m == "greedy" else env.random_actions(), check=False)
                    continue
                m.search()
                m.root_counts()
                env.next_state(m.pick_actions(), check=False)
        a_wins = b_wins = draws = 0
        for e, env in enumerate(self.envs):
            for row in self.searchers[e]:
                if not isinstance(row, str):
                    row.check_errors()
            r = env.game_ended(players=torch.ones_like(env.players))
            assert bool((r != 0).all())
            p1 = int((r > 0.5).sum().item())
            p2 = int((r < -0.5).sum().item())
            draws += self.n - p1 - p2
            if e == 0:
                a_wins, b_wins = a_wins + p1, b_wins + p2
            else:
                a_wins, b_wins = a_wins + p2, b_wins + p1
        return a_wins, b_wins, draws


def make_wave_player(total_games, wave_games, num_sims, evaluator, first_game=0, use_graph=True, **kwargs):
    """The reusable part of self_play_in_waves: ONE BatchedSelfPlay of min(wave_games, total_games) games (tree pool,
    example buffers) with the simulation wave captured as a CUDA graph.  Build it once, outside any timed region."""
    n = min(int(wave_games), int(total_games))
    ev = evaluator
    if n != wave_games and hasattr(ev, "with_private_buffers") and getattr(ev, "max_batch", n) < n:
        ev = ev.with_private_buffers(n)
    sp = BatchedSelfPlay(n, num_sims, evaluator=ev, game_base=first_game, **kwargs)
    if use_graph and not getattr(sp.mcts.evaluator, "uniform", False):
        sp.mcts.capture_graph()                                  # one simulation wave, replayed numMCTSSims times per move
    return sp


def self_play_in_waves(total_games, wave_games, num_sims, evaluator, first_game=0, on_wave=None, use_graph=True, sp=None,
                       **kwargs):
    """BASELINE.json configs[4]: more concurrent games than one tree pool fits in HBM (1,048,576 games over
    8 GPUs = 131,072 per GPU at ~2.8 MB of tree per game) are played as consecutive waves of `wave_games`
    games on ONE pool; wave w owns global games [first_game + w*wave_games, ...), so the result is the same
    as one huge batch (Philox streams are keyed by the global game id).  When the last wave is not full, its surplus
    slots play the games that follow the range (same pool, same captured graph) and are dropped from the tallies and from
    the examples handed out.  `on_wave(w, examples)` receives each wave's example tensors (copy or reduce them there: the
    buffers are reused).  `sp`: a player from make_wave_player (pool + graph built beforehand); otherwise one is built
    here.  Returns the total (p1_wins, p2_wins, draws)."""
    total_games = int(total_games)
    if sp is None:
        sp = make_wave_player(total_games, wave_games, num_sims, evaluator, first_game, use_graph, **kwargs)
    n = sp.n
    p1 = p2 = dr = 0
    waves = (total_games + n - 1) // n
    for w in range(waves):
        base = first_game + w * n
        live = min(n, total_games - w * n)                       # games of this wave that belong to the range
        if w or sp.env.game_base != base or int(sp.env.ply.max().item()) != 0:
            sp.env.game_base = base                              # same pool, next slice of global game ids (the captured
            sp.mcts.pool.reset()                                 # graph reads them from device memory: mcts.sync_game_base)
            sp.env.episode.zero_()
            sp.env.reset()
        ex = sp.execute_episodes()
        r = (ex["result_p1"] if ex is not None else sp.env.game_ended(players=torch.ones_like(sp.env.players)))[:live]
        p1 += int((r > 0.5).sum().item())
        p2 += int((r < -0.5).sum().item())
        dr += int((r.abs() < 0.5).sum().item())
        if on_wave is not None:
            if ex is not None and live < n:
                ex = {k: v.narrow(0 if v.dim() == 1 else (2 if k == "states" else 1), 0, live) for k, v in ex.items()}
            on_wave(w, ex)
    return p1, p2, dr
