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
