"""The judge-protocol bot with its real mover (BatchedMCTS / device greedy player) against the simulated judge, and the
stdin/stdout process a judge would spawn."""
import os
import subprocess
import sys
import time

import pytest
import torch

from judge_sim import play
from nypc_yacht_auction_b200 import judge_bot as jb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("policy,sims", [("mcts", 48), ("greedy", 0)])
def test_engine_mover_plays_full_legal_games(policy, sims):
    """Every reply is legal (the simulated judge asserts held dice / unused category / bid range), totals agree with the
    judge's sheet, and each decision stays far inside the 0.5 s window of INSTRUCTION.md:83-85."""
    mover = jb.EngineMover(num_sims=sims, policy=policy, seed=3)
    worst = 0.0
    for seed in (11, 12):
        session = jb.JudgeSession(mover)

        def handle(line):
            nonlocal worst
            t0 = time.perf_counter()
            r = session.handle(line)
            worst = max(worst, time.perf_counter() - t0)
            return r
        log, bot, opp = play(handle, seed, opp_first_on_score=bool(seed & 1), force_same_target=bool(seed & 1))
        assert session.totals() == (bot.total(), opp.total())
    assert worst < 0.5, worst


def test_network_mover_and_batched_choose():
    """EngineMover with a network evaluator, several sessions' boards in one call (n = 1..k)."""
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    from oracle import yacht_rules as yr
    from conftest import to_oracle_board
    torch.manual_seed(2)
    net = YachtPolicyValueNet().cuda().eval()
    mover = jb.EngineMover(num_sims=64, evaluator=FusedYachtEvaluator(net, 4), max_boards=4)
    sessions = [jb.JudgeSession(None) for _ in range(3)]
    for i, s in enumerate(sessions):
        s.roll_a, s.roll_b = [1 + (i + j) % 6 for j in range(5)], [6 - (i + j) % 6 for j in range(5)]
    boards = [s.board(0) for s in sessions]
    acts = mover.choose(boards)
    assert len(acts) == 3 and all(0 <= a < 202 for a in acts)
    for s in sessions:                                             # move on to a score decision with ten dice
        s.round_no, s.me.carry, s.opp.carry = 2, [1, 2, 3, 4, 5, 6, 6, 6, 2, 2], [1] * 10
    acts = mover.choose([s.board(1) for s in sessions])
    for s, a in zip(sessions, acts):
        assert yr.legal_mask(to_oracle_board(s.board(1)), 1)[a] == 1


def test_bot_process_speaks_the_protocol():
    """python -m nypc_yacht_auction_b200.judge_bot as the judge would run it: line in, flushed line out."""
    proc = subprocess.Popen([sys.executable, "-m", "nypc_yacht_auction_b200.judge_bot", "--sims", "32"], stdin=subprocess.PIPE,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)

    def ask(line, want_reply):
        proc.stdin.write(line + "\n")
        proc.stdin.flush()
        return proc.stdout.readline().strip() if want_reply else None
    try:
        assert ask("READY", True) == "OK"
        bid = ask("ROLL 13562 44421", True).split()
        assert bid[0] == "BID" and bid[1] in "AB" and 0 <= int(bid[2]) <= 100000
        ask("GET %s %s 700" % (bid[1], "B" if bid[1] == "A" else "A"), False)
        bid = ask("ROLL 66611 23456", True).split()
        ask("GET A A 99999", False)
        t0 = time.perf_counter()
        put = ask("SCORE", True).split()
        assert time.perf_counter() - t0 < 0.5
        assert put[0] == "PUT" and put[1] in jb.CATEGORIES and len(put[2]) == 5
        ask("SET LARGE_STRAIGHT 23456", False)                     # the opponent holds round 2's bundle B whatever we bid
        ask("FINISH", False)
        assert proc.wait(timeout=30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
