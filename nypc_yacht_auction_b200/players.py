"""Scripted players on the drop-in Game API (reference: yacht/YachtPlayers.py)."""
from __future__ import annotations

import numpy as np


class RandomYachtPlayer:
    """Uniform random legal move (yacht/YachtPlayers.py:174-183); consumes the global numpy RNG
    exactly like the reference so seeded arenas reproduce."""

    def __init__(self, game):
        self.game = game

    def play(self, board):
        valids = self.game.getValidMoves(board, 1)
        legal = np.nonzero(valids)[0]
        return int(np.random.choice(legal)) if len(legal) else 0


class GreedyYachtPlayer:
    """Heuristic bidder + greedy scorer (yacht/YachtPlayers.py:186-214) evaluated by ya_greedy_action;
    the random-legal fallback uses np.random.choice like the reference."""

    def __init__(self, game, seed=None):
        import random
        self.game = game
        if seed is not None:                     # yacht/YachtPlayers.py:192-196
            random.seed(seed)
            np.random.seed(seed)

    def play(self, board):
        import torch
        from . import _lib
        g = self.game
        g._put(board, 1)
        out = torch.zeros(1, dtype=torch.int32, device=g.device)
        _lib.check(g.lib.ya_greedy_action(_lib.ptr(g.d_state), 1, _lib.ptr(g.d_player), _lib.ptr(out), None, 1, 0, 0, 0,
                                          None, None, _lib.current_stream()), "ya_greedy_action")
        a = int(out.item())
        if a >= 0:
            return a
        valids = g.getValidMoves(board, 1)
        legal = np.nonzero(valids)[0]
        return int(np.random.choice(legal)) if len(legal) else 0
