"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure-Python restatement of the reference's Yacht-Auction rules (reference file
``yacht/YachtGame.py``; every function cites the lines it follows).  It is written
independently (histogram-based scoring, index-pair boards, explicit draw objects) and is
pinned against the reference itself by ``tests/golden/make_golden.py`` +
``tests/test_oracle_vs_golden.py`` and, when ``/root/reference`` is present, by the
side-by-side test ``tests/test_reference_interop.py``.

Small cases only (pure-Python loops).  The bulk oracle is ``oracle/yacht_oracle.c``.
"""
from __future__ import annotations

import itertools

import numpy as np

# yacht/YachtGame.py:15-50 -- constants and the action codec
N_CAT = 12
N_BID_LEVEL = 101
BID_UNIT = 500
N_BID = 2 * N_BID_LEVEL                      # 202
SUBSETS = tuple(itertools.combinations(range(10), 5))   # :35, lexicographic
N_SUBSET = len(SUBSETS)                      # 252
N_ACTION = N_BID + N_CAT * N_SUBSET          # 3226
BID, SCORE = 0, 1
LAST_ROUND = 13
UPPER_BONUS_AT = 63000
UPPER_BONUS = 35000


class Side:
    """One player's holdings (yacht/YachtGame.py:115-130)."""

    __slots__ = ("dice", "used", "cats", "bank")

    def __init__(self, dice=(), used=0, cats=None, bank=0):
        self.dice = list(dice)               # ordered carry, 0..10 dice
        self.used = used                     # 12-bit category mask
        self.cats = list(cats) if cats is not None else [0] * N_CAT
        self.bank = bank                     # net auction result

    def clone(self):
        return Side(self.dice, self.used, self.cats, self.bank)

    def total(self):
        # yacht/YachtGame.py:125-130
        upper = sum(self.cats[:6])
        return sum(self.cats) + (UPPER_BONUS if upper >= UPPER_BONUS_AT else 0) + self.bank

    def finished(self):
        # yacht/YachtGame.py:189-190
        return self.used == (1 << N_CAT) - 1


class Board:
    """Full game state (yacht/YachtGame.py:133-145).  sides[0] is "p1", sides[1] is "p2"."""

    __slots__ = ("rnd", "phase", "pool_a", "pool_b", "bids", "sides")

    def __init__(self):
        self.rnd = 1
        self.phase = BID
        self.pool_a = []
        self.pool_b = []
        self.bids = [None, None]             # (target 0=A/1=B, amount) or None
        self.sides = [Side(), Side()]

    def clone(self):
        # yacht/YachtGame.py:480-500
        b = Board()
        b.rnd, b.phase = self.rnd, self.phase
        b.pool_a, b.pool_b = list(self.pool_a), list(self.pool_b)
        b.bids = list(self.bids)
        b.sides = [self.sides[0].clone(), self.sides[1].clone()]
        return b


def category_points(cat, five):
    """yacht/YachtGame.py:57-108 via a face histogram."""
    hist = [0] * 7
    for d in five:
        hist[d] += 1
    pips = sum(five)
    if cat < 6:
        return 1000 * (cat + 1) * hist[cat + 1]
    if cat == 6:
        return 1000 * pips
    if cat == 7:
        return 1000 * pips if max(hist[1:]) >= 4 else 0
    if cat == 8:
        has2 = any(h in (2, 5) for h in hist[1:])
        has3 = any(h in (3, 5) for h in hist[1:])
        return 1000 * pips if (has2 and has3) else 0
    seen = sum(1 << (f - 1) for f in range(1, 7) if hist[f])
    if cat == 9:
        return 15000 if any((seen & m) == m for m in (0b001111, 0b011110, 0b111100)) else 0
    if cat == 10:
        return 30000 if any((seen & m) == m for m in (0b011111, 0b111110)) else 0
    if cat == 11:
        return 50000 if max(hist[1:]) == 5 else 0
    raise ValueError("Invalid category")


def new_game(draw):
    """yacht/YachtGame.py:232-237: roll A then B."""
    b = Board()
    b.pool_a = draw.roll_a()
    b.pool_b = draw.roll_b()
    return b


def _settle_auction(b, draw):
    """yacht/YachtGame.py:502-542."""
    (t0, a0), (t1, a1) = b.bids
    got = [t0, t1]
    if t0 == t1:
        if a0 > a1:
            win = 0
        elif a1 > a0:
            win = 1
        else:
            win = draw.tie()                 # :522 -- only drawn on an exact tie
        got[1 - win] = 1 - got[win]
    b.sides[0].bank += -a0 if got[0] == t0 else a0
    b.sides[1].bank += -a1 if got[1] == t1 else a1
    pools = (b.pool_a, b.pool_b)
    b.sides[0].dice.extend(pools[got[0]])
    b.sides[1].dice.extend(pools[got[1]])


def next_state(board, player, action, draw):
    """yacht/YachtGame.py:260-372.  ``draw`` supplies tie() / roll_a() / roll_b() in that order."""
    b = board.clone()
    me = 0 if player == 1 else 1
    if b.phase == BID and b.rnd != LAST_ROUND:
        if not 0 <= action < N_BID:
            raise ValueError("Invalid action in BID phase")          # :268-269
        bid = (action // N_BID_LEVEL, (action % N_BID_LEVEL) * BID_UNIT)
        first = b.bids[0] is None and b.bids[1] is None
        b.bids[me] = bid
        if first:
            return b, -player                                         # :272-279
        assert b.bids[0] is not None and b.bids[1] is not None       # :508
        _settle_auction(b, draw)
        if b.rnd != 1:
            b.phase = SCORE                                           # :290-293
        else:
            b.rnd += 1                                                # :295-301
            b.bids = [None, None]
            b.pool_a = draw.roll_a()
            b.pool_b = draw.roll_b()
        return b, 1
    if b.phase == SCORE:
        if not N_BID <= action < N_ACTION:
            raise ValueError("Invalid action in SCORE phase")        # :306-307
        cat, sub = divmod(action - N_BID, N_SUBSET)
        pos = SUBSETS[sub]
        side = b.sides[me]
        if (side.used >> cat) & 1 or pos[-1] >= len(side.dice):
            return b, -player                                         # :312-324 silent no-op
        side.cats[cat] = category_points(cat, [side.dice[i] for i in pos])
        side.dice = [d for i, d in enumerate(side.dice) if i not in pos]
        side.used |= 1 << cat
        if b.rnd == LAST_ROUND:                                       # :338-349
            if b.sides[0].finished() and b.sides[1].finished():
                return b, 1
            return b, -player
        if player == -1:                                              # :352-365
            b.rnd += 1
            b.bids = [None, None]
            if b.rnd != LAST_ROUND:
                b.pool_a = draw.roll_a()
                b.pool_b = draw.roll_b()
                b.phase = BID
            else:
                b.phase = SCORE
            return b, 1
        return b, -player                                             # :366-369
    raise RuntimeError("Invalid phase/state")                         # :372


def legal_count(board, player):
    """Number of ones in legal_mask (closed form of yacht/YachtGame.py:374-406)."""
    if board.phase == BID and board.rnd != LAST_ROUND:
        return N_BID
    if board.phase == SCORE:
        side = board.sides[0 if player == 1 else 1]
        n = len(side.dice)
        if n < 5:
            return 0
        free = N_CAT - bin(side.used).count("1")
        return free * sum(1 for s in SUBSETS if s[-1] < n)
    return 0


def legal_mask(board, player):
    """yacht/YachtGame.py:374-406."""
    v = np.zeros(N_ACTION, dtype=np.uint8)
    if board.phase == BID and board.rnd != LAST_ROUND:
        v[:N_BID] = 1
        return v
    if board.phase == SCORE:
        side = board.sides[0 if player == 1 else 1]
        n = len(side.dice)
        if n < 5:
            return v
        ok = np.array([s[-1] < n for s in SUBSETS], dtype=np.uint8)
        for cat in range(N_CAT):
            if not (side.used >> cat) & 1:
                v[N_BID + cat * N_SUBSET: N_BID + (cat + 1) * N_SUBSET] = ok
    return v


def outcome(board, player):
    """yacht/YachtGame.py:408-428."""
    if not (board.sides[0].finished() and board.sides[1].finished()):
        return 0.0
    t0, t1 = board.sides[0].total(), board.sides[1].total()
    if t0 == t1:
        return 1e-4
    lead = 1 if t0 > t1 else -1
    return float(lead if player == 1 else -lead)


def canonical(board, player):
    """yacht/YachtGame.py:430-442 (player 1 aliases the input, as the reference does)."""
    if player == 1:
        return board
    b = board.clone()
    b.sides.reverse()
    b.bids.reverse()
    return b


def key(board):
    """yacht/YachtGame.py:448-467 -- byte-identical text key."""
    def digits(d):
        return "".join(str(int(x)) for x in d)

    def bid(x):
        return "-" if x is None else "AB"[x[0]] + str(x[1])

    s0, s1 = board.sides
    parts = [
        "r%d" % board.rnd, "ph%d" % board.phase,
        "A" + (digits(board.pool_a) if board.pool_a else "-"),
        "B" + (digits(board.pool_b) if board.pool_b else "-"),
        "p1b" + bid(board.bids[0]), "p2b" + bid(board.bids[1]),
        "p1c" + digits(s0.dice), "p2c" + digits(s1.dice),
        "p1u%d" % s0.used, "p2u%d" % s1.used,
        "p1s" + ",".join(str(x) for x in s0.cats), "p2s" + ",".join(str(x) for x in s1.cats),
        "p1bid%d" % s0.bank, "p2bid%d" % s1.bank,
    ]
    return "|".join(parts)


def features(board):
    """yacht/NNet.py:50-86 (state_to_vec) for a canonical board -> float32[59]."""
    def die(d):
        return (d - 3.5) / 3.5

    def padded(dice, n):
        out = np.full(n, -1.0, dtype=np.float32)
        for i, d in enumerate(dice[:n]):
            out[i] = die(d)
        return out

    show = board.phase == BID and board.rnd != LAST_ROUND
    vec = [board.rnd / 13.0, 1.0 if board.phase == BID else 0.0, 1.0 if board.phase == SCORE else 0.0]
    vec.extend(padded(board.sides[0].dice, 10))
    vec.extend(padded(board.sides[1].dice, 10))
    vec.extend(padded(board.pool_a if show else [], 5))
    vec.extend(padded(board.pool_b if show else [], 5))
    for side in board.sides:
        vec.extend(np.array([(side.used >> i) & 1 for i in range(N_CAT)], dtype=np.float32))
    vec.append(board.sides[0].bank * 1e-5)
    vec.append(board.sides[1].bank * 1e-5)
    return np.asarray(vec, dtype=np.float32)


def nth_legal_action(board, player, idx):
    """The idx-th set position of legal_mask, in closed form (used by the random policy)."""
    if board.phase == BID and board.rnd != LAST_ROUND:
        return idx
    side = board.sides[0 if player == 1 else 1]
    n = len(side.dice)
    per_cat = sum(1 for s in SUBSETS if s[-1] < n)
    k, sub = divmod(idx, per_cat)
    free = [c for c in range(N_CAT) if not (side.used >> c) & 1]
    return N_BID + free[k] * N_SUBSET + sub


def random_legal_action(board, player, draw):
    """yacht/YachtPlayers.py:174-183 semantics (uniform over the legal set, 0 if empty), with
    the engine's Philox pick instead of np.random.choice."""
    count = legal_count(board, player)
    if count == 0:
        return 0
    return nth_legal_action(board, player, draw.pick(count))


def score_table(dice, used=0):
    """All 12 x 252 (category, subset) scores / 1000 for an ordered carry, 0 where the subset
    does not fit (the enumeration of yacht/YachtPlayers.py:134-169 without the argmax)."""
    out = np.zeros((N_CAT, N_SUBSET), dtype=np.uint8)
    n = len(dice)
    for ci, pos in enumerate(SUBSETS):
        if pos[-1] >= n:
            continue
        five = [dice[i] for i in pos]
        for cat in range(N_CAT):
            out[cat, ci] = category_points(cat, five) // 1000
    return out
