"""Host mirror of the packed 32-byte game state (bit layout: csrc/ya_common.cuh).

``YachtBoard`` is the board object the drop-in ``YachtGame`` hands to Coach / Arena / MCTS /
``state_to_vec``: it carries the 8 packed words and decodes, on attribute access, the fields
the reference's ``YachtState`` / ``PlayerState`` dataclasses expose
(/root/reference/yacht/YachtGame.py:115-145), so code written against those attributes
(yacht/NNet.py:65-86, Arena.py:78-84, yacht/YachtPlayers.py) runs unchanged.

This module only (de)serialises; no game rule is evaluated on the host.
"""
from __future__ import annotations

import numpy as np

N_CAT = 12
WORDS = 8


class PlayerView:
    """Decoded per-player fields (reference PlayerState, YachtGame.py:115-130)."""

    __slots__ = ("carry", "used_mask", "cat_scores", "bid_score")

    def __init__(self, carry, used_mask, cat_scores, bid_score):
        self.carry = carry
        self.used_mask = used_mask
        self.cat_scores = cat_scores
        self.bid_score = bid_score

    def basic_total(self):
        return sum(self.cat_scores[0:6])

    def total_with_bonus(self):
        return sum(self.cat_scores) + (35000 if self.basic_total() >= 63000 else 0) + self.bid_score

    def __repr__(self):
        return "PlayerView(carry=%r, used_mask=%d, cat_scores=%r, bid_score=%d)" % (
            self.carry, self.used_mask, self.cat_scores, self.bid_score)


def _dice(word, count):
    out = []
    for i in range(count):
        d = (word >> (3 * i)) & 7
        if d == 0:
            break
        out.append(d)
    return out


def _bid(slot):
    if not slot & 1:
        return None
    return ("B" if (slot >> 1) & 1 else "A", (slot >> 2) * 500)


def _player(w_carry, w4, w5):
    cats = [1000 * (c + 1) * ((w5 >> (3 * c)) & 7) for c in range(6)]
    cats.append(1000 * ((w5 >> 18) & 31))
    cats.append(1000 * ((w5 >> 23) & 31))
    cats.append(1000 * ((w4 >> 25) & 31))
    cats.append(15000 * ((w5 >> 28) & 1))
    cats.append(30000 * ((w5 >> 29) & 1))
    cats.append(50000 * ((w5 >> 30) & 1))
    bank = (w4 >> 12) & 0x1FFF
    if bank >= 0x1000:
        bank -= 0x2000
    return PlayerView(_dice(w_carry, 10), w4 & 0xFFF, cats, bank * 500)


class YachtBoard:
    """Value object: 8 packed words + lazily decoded reference-style attributes."""

    __slots__ = ("words", "_p1", "_p2")

    def __init__(self, words):
        self.words = tuple(int(x) & 0xFFFFFFFF for x in words)
        assert len(self.words) == WORDS
        self._p1 = None
        self._p2 = None

    # -- reference YachtState fields -------------------------------------------------------
    @property
    def round_no(self):
        return self.words[0] & 15

    @property
    def phase(self):
        return (self.words[0] >> 4) & 1

    @property
    def rollA(self):
        return _dice(self.words[1] & 0x7FFF, 5)

    @property
    def rollB(self):
        return _dice((self.words[1] >> 15) & 0x7FFF, 5)

    @property
    def p1_bid(self):
        return _bid((self.words[0] >> 5) & 0x1FF)

    @property
    def p2_bid(self):
        return _bid((self.words[0] >> 14) & 0x1FF)

    @property
    def p1(self):
        if self._p1 is None:
            self._p1 = _player(self.words[2], self.words[4], self.words[5])
        return self._p1

    @property
    def p2(self):
        if self._p2 is None:
            self._p2 = _player(self.words[3], self.words[6], self.words[7])
        return self._p2

    # -- value semantics -------------------------------------------------------------------
    def __eq__(self, other):
        return isinstance(other, YachtBoard) and self.words == other.words

    def __hash__(self):
        return hash(self.words)

    def __reduce__(self):
        return (YachtBoard, (self.words,))

    def __repr__(self):
        return "YachtBoard(%s)" % string_key(self)

    def to_numpy(self):
        return np.asarray(self.words, dtype=np.uint32)


def string_key(b):
    """Text key in the reference's stringRepresentation format (YachtGame.py:448-467); works on a
    YachtBoard or on any object with the reference's attributes."""
    def digits(d):
        return "".join(str(int(x)) for x in d)

    def bid(x):
        return "-" if x is None else "%s%d" % (x[0], x[1])

    return "|".join([
        "r%d" % b.round_no, "ph%d" % b.phase,
        "A" + (digits(b.rollA) if len(b.rollA) else "-"), "B" + (digits(b.rollB) if len(b.rollB) else "-"),
        "p1b" + bid(b.p1_bid), "p2b" + bid(b.p2_bid),
        "p1c" + digits(b.p1.carry), "p2c" + digits(b.p2.carry),
        "p1u%d" % b.p1.used_mask, "p2u%d" % b.p2.used_mask,
        "p1s" + ",".join(str(int(x)) for x in b.p1.cat_scores), "p2s" + ",".join(str(int(x)) for x in b.p2.cat_scores),
        "p1bid%d" % b.p1.bid_score, "p2bid%d" % b.p2.bid_score,
    ])


def _pack_dice(dice, limit, what):
    dice = [int(d) for d in dice]
    if len(dice) > limit:
        raise ValueError("%s holds %d dice; the packed state supports at most %d" % (what, len(dice), limit))
    w = 0
    for i, d in enumerate(dice):
        if not 1 <= d <= 6:
            raise ValueError("%s: die value %r out of range" % (what, d))
        w |= d << (3 * i)
    return w


def _pack_bid(b, what):
    if b is None:
        return 0
    target, amount = b
    amount = int(amount)
    if target not in ("A", "B") or amount % 500 or not 0 <= amount <= 50000:
        raise ValueError("%s: bid %r is not on the reference's 0..50000 step 500 grid" % (what, b))
    return 1 | ((1 if target == "B" else 0) << 1) | ((amount // 500) << 2)


def _pack_player(p, what):
    used = int(p.used_mask)
    if not 0 <= used < 4096:
        raise ValueError("%s: used_mask out of range" % what)
    cats = [int(x) for x in p.cat_scores]
    if len(cats) != N_CAT:
        raise ValueError("%s: cat_scores must have 12 entries" % what)
    w4, w5 = used, 0
    for c in range(6):
        unit = 1000 * (c + 1)
        if cats[c] % unit or not 0 <= cats[c] // unit <= 5:
            raise ValueError("%s: cat_scores[%d]=%d is not a reachable score" % (what, c, cats[c]))
        w5 |= (cats[c] // unit) << (3 * c)
    for c, shift, word in ((6, 18, 5), (7, 23, 5), (8, 25, 4)):
        if cats[c] % 1000 or not 0 <= cats[c] // 1000 <= 30:
            raise ValueError("%s: cat_scores[%d]=%d is not a reachable score" % (what, c, cats[c]))
        if word == 5:
            w5 |= (cats[c] // 1000) << shift
        else:
            w4 |= (cats[c] // 1000) << shift
    for c, full, bit in ((9, 15000, 28), (10, 30000, 29), (11, 50000, 30)):
        if cats[c] not in (0, full):
            raise ValueError("%s: cat_scores[%d]=%d is not a reachable score" % (what, c, cats[c]))
        w5 |= (1 if cats[c] else 0) << bit
    bank = int(p.bid_score)
    if bank % 500 or not -4096 <= bank // 500 <= 4095:
        raise ValueError("%s: bid_score %d not representable" % (what, bank))
    w4 |= ((bank // 500) & 0x1FFF) << 12
    return _pack_dice(p.carry, 10, what + ".carry"), w4, w5


def pack_state(s):
    """Pack any object exposing the reference's YachtState attributes into a YachtBoard."""
    if isinstance(s, YachtBoard):
        return s
    rnd, phase = int(s.round_no), int(s.phase)
    if not 1 <= rnd <= 13 or phase not in (0, 1):
        raise ValueError("round/phase out of range")
    w0 = rnd | (phase << 4) | (_pack_bid(s.p1_bid, "p1_bid") << 5) | (_pack_bid(s.p2_bid, "p2_bid") << 14)
    w1 = _pack_dice(s.rollA, 5, "rollA") | (_pack_dice(s.rollB, 5, "rollB") << 15)
    c1, w4, w5 = _pack_player(s.p1, "p1")
    c2, w6, w7 = _pack_player(s.p2, "p2")
    return YachtBoard((w0, w1, c1, c2, w4, w5, w6, w7))


def boards_to_planes(boards):
    """List of YachtBoard -> uint32[2, n, 4] in the device plane layout."""
    arr = np.asarray([b.words for b in boards], dtype=np.uint32).reshape(len(boards), 2, 4)
    return np.ascontiguousarray(arr.transpose(1, 0, 2))


def planes_to_boards(planes):
    """uint32[2, n, 4] -> list of YachtBoard."""
    arr = np.asarray(planes).astype(np.uint32, copy=False)
    n = arr.shape[1]
    rows = arr.transpose(1, 0, 2).reshape(n, 8)
    return [YachtBoard(r) for r in rows]
