"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure-Python Philox4x32-10 and the engine's *draw protocol*, restated independently of the
CUDA code so the oracle can replay exactly the dice / tie-breaks / sampled actions the
device produces (SURVEY.md section 8c "dice injection protocol").

The reference has no counter-based RNG: it draws dice with ``np.random.randint(1, 7, 5)``
(yacht/YachtGame.py:154-155) and tie-breaks with ``random.randint(0, 1)``
(yacht/YachtGame.py:158-159).  Those two hooks are the injection points; this module
defines what gets injected.

Draw protocol (one Philox block = one draw event)::

    key  = (seed & 0xffffffff, seed >> 32)
    ctr  = (game_id, episode, ply | tag << 8 | depth << 16, sim)
    w0..w3 = philox4x32_10(ctr, key)
    rollA  = five_dice(w0)        rollB = five_dice(w1)
    tie    = w2 >> 31             pick  = (w3 * L) >> 32     # uniform index in [0, L)

    five_dice(w): repeat 5x:  t = w * 6;  die = 1 + (t >> 32);  w = t & 0xffffffff
"""
from __future__ import annotations

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

TAG_INIT = 0      # getInitBoard rolls (ply = 0)
TAG_REAL = 1      # real-game getNextState at ply p
TAG_ACTION = 2    # policy sampling at ply p
TAG_SEARCH = 3    # getNextState inside MCTS.search: (ply, sim, depth)


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def five_dice(w):
    out = []
    for _ in range(5):
        t = w * 6
        out.append(1 + (t >> 32))
        w = t & MASK
    return out


def draw_words(seed, game_id, episode, ply, tag, depth=0, sim=0):
    key = (seed & MASK, (seed >> 32) & MASK)
    ctr = (game_id & MASK, episode & MASK, (ply & 0xFF) | ((tag & 0xFF) << 8) | ((depth & 0xFF) << 16), sim & MASK)
    return philox4x32_10(ctr, key)


class Draw:
    """One draw event: exposes the pieces a transition may consume."""

    __slots__ = ("w",)

    def __init__(self, seed, game_id, episode, ply, tag, depth=0, sim=0):
        self.w = draw_words(seed, game_id, episode, ply, tag, depth, sim)

    def roll_a(self):
        return five_dice(self.w[0])

    def roll_b(self):
        return five_dice(self.w[1])

    def tie(self):
        return self.w[2] >> 31

    def pick(self, count):
        return (self.w[3] * count) >> 32
