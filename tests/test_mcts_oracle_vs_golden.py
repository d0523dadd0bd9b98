"""Pins oracle/mcts_oracle.py against traces of the reference's own MCTS.py
(tests/golden/mcts_golden.json, made by tests/golden/make_golden_mcts.py)."""
import json
import os

import numpy as np
import pytest

from oracle import mcts_oracle, philox
from oracle import yacht_rules as yr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_golden.json")

EVALUATORS = {
    "uniform_s25": mcts_oracle.uniform_evaluator,
    "hashed_s30": mcts_oracle.hashed_evaluator,
    "hashed_s64_temp0": lambda b: mcts_oracle.hashed_evaluator(b, 7),
    "hashed_s200_late": lambda b: mcts_oracle.hashed_evaluator(b, 3),
}


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def pairwise_sum_f32(a):
    """numpy's float32 add.reduce order (pairwise, 8 accumulators, blocks of <= 128)."""
    f = np.float32
    n = len(a)
    if n < 8:
        res = f(0.0)
        for x in a:
            res = f(res + x)
        return res
    if n <= 128:
        r = [f(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def test_numpy_sum_is_the_documented_pairwise_order():
    """The CUDA expand kernel re-implements this order; make sure this machine's numpy uses it."""
    rng = np.random.default_rng(1)
    for n in (3226, 3226, 3226, 202, 3024, 129, 128, 64, 9, 7):
        a = rng.random(n, dtype=np.float32)
        a[rng.random(n) < 0.4] = 0
        assert pairwise_sum_f32(a) == np.sum(a)


@pytest.mark.parametrize("case", load_cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_mcts(case):
    trace, board, cur, r, tree = mcts_oracle.self_play_game(
        EVALUATORS[case["name"]], case["sims"], case["cpuct"], case["seed"], case["game"],
        temp_threshold=case["temp_threshold"], max_plies=case["max_plies"])
    assert len(trace) == len(case["trace"])
    for mine, ref in zip(trace, case["trace"]):
        assert mine["key"] == ref["key"], mine["ply"]
        assert {str(k): v for k, v in mine["counts"].items()} == ref["counts"], mine["ply"]
        assert mine["action"] == ref["action"]
        assert mine["nodes"] == ref["nodes"]
        node = tree.nodes[ref["key"]] if ref is case["trace"][-1] else None
    assert yr.key(board) == case["final_key"]
    assert float(r) == case["result"] and cur == case["final_player"]
    assert tree.leaf_evals == case["leaf_evals"] and len(tree.nodes) == case["total_nodes"]
    assert len(tree.terminal) == case["terminal_states"]
    # root Q values of the last searched ply, including their numeric type (float32 vs double)
    last = case["trace"][-1]
    node = tree.nodes[last["key"]]
    for a, (kind, hexval) in last["q"].items():
        q = node.q_edge[int(a)]
        assert ("f32" if isinstance(q, np.float32) else "f64") == kind
        assert float(q).hex() == hexval


def test_survey_known_answers():
    """uniform evaluator, 25 sims, cpuct 1.5: the 24 lowest legal actions get one
    visit each on a fresh root; a reused second-bidder root has Ns = 25."""
    case = [c for c in load_cases() if c["name"] == "uniform_s25"][0]
    t = case["trace"]
    assert t[0]["counts"] == {str(a): 1 for a in range(24)} and t[0]["ns"] == 24
    assert t[1]["ns"] == 25
    assert t[-1]["counts"] == {"2974": 48} or list(t[-1]["counts"].values()) == [48]
