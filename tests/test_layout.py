"""Host-side packing: pack -> unpack is the identity on the reference's own states (golden keys),
and rejects states the 32-byte layout cannot hold."""
import json
import re

import pytest
from hypothesis import given, settings, strategies as st

from nypc_yacht_auction_b200.layout import YachtBoard, pack_state, string_key, boards_to_planes, planes_to_boards


class _P:
    def __init__(self, carry, used, cats, bank):
        self.carry, self.used_mask, self.cat_scores, self.bid_score = carry, used, cats, bank


class _S:
    pass


def parse_key(key):
    f = dict()
    parts = key.split("|")
    s = _S()
    s.round_no = int(parts[0][1:])
    s.phase = int(parts[1][2:])
    s.rollA = [] if parts[2] == "A-" else [int(c) for c in parts[2][1:]]
    s.rollB = [] if parts[3] == "B-" else [int(c) for c in parts[3][1:]]

    def bid(x):
        x = x[3:]
        return None if x == "-" else (x[0], int(x[1:]))
    s.p1_bid, s.p2_bid = bid(parts[4]), bid(parts[5])
    c1, c2 = [int(c) for c in parts[6][3:]], [int(c) for c in parts[7][3:]]
    u1, u2 = int(parts[8][3:]), int(parts[9][3:])
    s1, s2 = [int(x) for x in parts[10][3:].split(",")], [int(x) for x in parts[11][3:].split(",")]
    b1, b2 = int(parts[12][5:]), int(parts[13][5:])
    s.p1, s.p2 = _P(c1, u1, s1, b1), _P(c2, u2, s2, b2)
    return s


def test_roundtrip_on_reference_states(rules_golden):
    meta, _ = rules_golden
    keys = set()
    for tr in meta["seeded"]:
        for ply in tr["plies"]:
            keys.add(ply["canon_key"])
            keys.add(ply["next_key"])
    assert len(keys) > 250
    boards = []
    for k in keys:
        b = pack_state(parse_key(k))
        assert string_key(b) == k
        assert pack_state(b) is b
        boards.append(b)
    assert len(set(boards)) == len(keys)                 # injective
    back = planes_to_boards(boards_to_planes(boards))
    assert back == boards
    import pickle
    assert pickle.loads(pickle.dumps(boards[0])) == boards[0]


def test_rejects_unrepresentable_states():
    s = parse_key("r5|ph1|A11111|B22222|p1b-|p2b-|p1c12345|p2c|p1u0|p2u0|p1s0,0,0,0,0,0,0,0,0,0,0,0|"
                  "p2s0,0,0,0,0,0,0,0,0,0,0,0|p1bid0|p2bid0")
    pack_state(s)
    s.p1.carry = [1] * 11
    with pytest.raises(ValueError):
        pack_state(s)
    s.p1.carry = [7]
    with pytest.raises(ValueError):
        pack_state(s)
    s.p1.carry = []
    s.p1.cat_scores[1] = 3000          # TWO can only score multiples of 2000
    with pytest.raises(ValueError):
        pack_state(s)
    s.p1.cat_scores[1] = 0
    s.p1_bid = ("A", 250)
    with pytest.raises(ValueError):
        pack_state(s)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 13), st.integers(0, 1), st.lists(st.integers(1, 6), min_size=0, max_size=10),
       st.integers(0, 4095), st.integers(-1200, 1200), st.lists(st.integers(0, 5), min_size=6, max_size=6),
       st.lists(st.integers(0, 30), min_size=3, max_size=3), st.lists(st.booleans(), min_size=3, max_size=3),
       st.one_of(st.none(), st.tuples(st.sampled_from("AB"), st.integers(0, 100).map(lambda x: x * 500))))
def test_roundtrip_property(rnd, phase, carry, used, bank, upper, mids, flags, bid):
    s = _S()
    s.round_no, s.phase, s.rollA, s.rollB, s.p1_bid, s.p2_bid = rnd, phase, [1, 2, 3, 4, 5], [6, 5, 4, 3, 2], bid, None
    cats = [1000 * (c + 1) * upper[c] for c in range(6)] + [1000 * m for m in mids] + \
           [15000 * flags[0], 30000 * flags[1], 50000 * flags[2]]
    s.p1 = _P(list(carry), used, cats, bank * 500)
    s.p2 = _P([], 0, [0] * 12, -bank * 500)
    b = pack_state(s)
    assert b.round_no == rnd and b.phase == phase and b.p1_bid == bid and b.p2_bid is None
    assert b.p1.carry == list(carry) and b.p1.used_mask == used and b.p1.cat_scores == cats
    assert b.p1.bid_score == bank * 500 and b.p2.bid_score == -bank * 500
    assert b.rollA == [1, 2, 3, 4, 5] and b.rollB == [6, 5, 4, 3, 2]
    assert YachtBoard(b.words) == b
