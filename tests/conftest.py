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
