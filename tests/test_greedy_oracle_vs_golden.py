"""Pins oracle/greedy_oracle.py against the reference's GreedyYachtPlayer (tests/golden/greedy_golden.json)."""
import json
import os

from oracle import greedy_oracle
from conftest import to_oracle_board
from test_layout import parse_key

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "greedy_golden.json")


def test_greedy_choices_match_reference():
    with open(GOLDEN) as f:
        games = json.load(f)["games"]
    checked = bids = 0
    for gm in games:
        for p in gm["plies"]:
            if p["kind"] != "greedy":
                continue
            board = to_oracle_board(parse_key(p["key"]))
            a, legal = greedy_oracle.greedy_action(board)
            assert a == p["raw"], p["key"]
            assert legal == (p["raw"] == p["action"])
            checked += 1
            bids += p["raw"] < 202
    assert checked >= 200 and bids >= 90


def test_bid_overflow_quirk_q11():
    """Bids are clipped to 100000 (not 50000): a huge deficit drives the encoded action out of the A range."""
    b = to_oracle_board(parse_key(
        "r5|ph0|A66666|B12345|p1b-|p2b-|p1c11234|p2c23456|p1u7|p2u7|p1s0,0,0,0,0,0,0,0,0,0,0,0|"
        "p2s5000,10000,15000,0,0,0,0,0,0,0,0,0|p1bid-400000|p2bid300000"))
    a, legal = greedy_oracle.greedy_action(b)
    assert a >= 101                      # amount index > 100
