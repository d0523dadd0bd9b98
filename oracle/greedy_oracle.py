"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the reference's heuristic player (/root/reference/yacht/YachtPlayers.py:39-214):
greedy best-immediate-gain scoring and the value-gap bid heuristic, including its quirks (first
maximum wins; bids are clipped to 100000, not 50000, so the encoded action can alias into the B range
or leave the bid range, YachtPlayers.py:126-129 -- SURVEY.md quirk Q11).  Pinned by
tests/golden/greedy_golden.json (made from the unmodified reference).
"""
from __future__ import annotations

from . import yacht_rules as yr


def _gain(cat, five, upper_before):
    sc = yr.category_points(cat, five)
    if cat < 6 and upper_before < yr.UPPER_BONUS_AT <= upper_before + sc:
        return sc + yr.UPPER_BONUS                       # YachtPlayers.py:88-91 / 158-162
    return sc


def best_gain(side, dice):
    """max over unused categories x fitting subsets of the immediate gain; (gain, action) with the first
    maximum in (category, subset) order, or (None, None) if nothing is playable."""
    n = len(dice)
    upper_before = sum(side.cats[:6])
    best, best_action = None, None
    for cat in range(yr.N_CAT):
        if (side.used >> cat) & 1:
            continue
        for ci, pos in enumerate(yr.SUBSETS):
            if pos[-1] >= n:
                break                                    # later subsets do not fit either (:153-154)
            g = _gain(cat, [dice[i] for i in pos], upper_before)
            if best is None or g > best:
                best, best_action = g, yr.N_BID + cat * yr.N_SUBSET + ci
    return best, best_action


def bundle_value(board, side, bundle):
    """YachtPlayers.py:39-95."""
    if board.rnd == 1:
        dice = side.dice + bundle
        s = 1000 * sum(dice)
        counts = [dice.count(v) for v in range(1, 7)]
        if max(counts) >= 4:
            s += 6000
        elif max(counts) == 3:
            s += 3000
        e = [c > 0 for c in counts]
        if (e[0] and e[1] and e[2] and e[3]) or (e[1] and e[2] and e[3] and e[4]) or (e[2] and e[3] and e[4] and e[5]):
            s += 5000
        return s
    dice = side.dice + bundle
    if len(dice) < 5:
        return 0
    best, _ = best_gain(side, dice)
    return 0 if best is None else best


def choose_bid(board):
    """YachtPlayers.py:98-129 on a canonical board."""
    me, opp = board.sides
    val_a = bundle_value(board, me, board.pool_a)
    val_b = bundle_value(board, me, board.pool_b)
    if val_a >= val_b:
        target, gap = 0, max(0, val_a - val_b)
    else:
        target, gap = 1, max(0, val_b - val_a)
    diff = me.total() - opp.total()
    bid_k = 0.5 * (gap / 1000.0) - 0.15 * (diff / 1000.0)
    bid = int(max(0, min(100000, round(1000 * bid_k))))
    bid = (bid // yr.BID_UNIT) * yr.BID_UNIT
    return target * yr.N_BID_LEVEL + bid // yr.BID_UNIT


def choose_scoring(board):
    """YachtPlayers.py:134-169."""
    me = board.sides[0]
    if len(me.dice) < 5:
        return 0
    _, action = best_gain(me, me.dice)
    return action if action is not None else 0


def greedy_action(board):
    """GreedyYachtPlayer.play without the random fallback: returns (action, is_legal)."""
    mask = yr.legal_mask(board, 1)
    if board.phase == yr.BID and board.rnd != yr.LAST_ROUND:
        a = choose_bid(board)
    else:
        a = choose_scoring(board)
    return a, bool(a < yr.N_ACTION and mask[a] == 1)
