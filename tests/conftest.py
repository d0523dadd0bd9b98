import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rules_golden():
    with open(os.path.join(GOLDEN, "rules_golden.json")) as f:
        meta = json.load(f)
    arrays = dict(np.load(os.path.join(GOLDEN, "rules_golden.npz")))
    return meta, arrays


class TapeDraw:
    """Replays recorded reference draws ([["roll", [..]] | ["tie", t], ...]) in consumption order."""

    def __init__(self, tape):
        self.tape = list(tape)

    def _pop(self, kind):
        k, v = self.tape.pop(0)
        assert k == kind, (k, kind)
        return v

    def roll_a(self):
        return list(self._pop("roll"))

    def roll_b(self):
        return list(self._pop("roll"))

    def tie(self):
        return self._pop("tie")

    def done(self):
        return not self.tape


def to_oracle_board(yb):
    """YachtBoard (or anything with the reference's attributes) -> oracle.yacht_rules.Board."""
    from oracle import yacht_rules as yr
    b = yr.Board()
    b.rnd, b.phase = yb.round_no, yb.phase
    b.pool_a, b.pool_b = [int(x) for x in yb.rollA], [int(x) for x in yb.rollB]

    def conv(x):
        return None if x is None else ("AB".index(x[0]), int(x[1]))
    b.bids = [conv(yb.p1_bid), conv(yb.p2_bid)]
    for i, p in enumerate((yb.p1, yb.p2)):
        b.sides[i] = yr.Side([int(x) for x in p.carry], int(p.used_mask), [int(x) for x in p.cat_scores], int(p.bid_score))
    return b
