"""Judge-protocol client (nypc_yacht_auction_b200/judge_bot.py) on the CPU: the protocol state machine against a scripted
transcript in the format of INSTRUCTION.md:76-92, full simulated games with a stub mover (real-valued opponent bids), and
its score-sheet arithmetic against the oracle on all 7776 x 12 inputs.  The GPU mover is covered by test_judge_bot_gpu.py."""
import itertools

import numpy as np
import pytest

from oracle import yacht_rules as yr
from conftest import to_oracle_board
from judge_sim import play
from nypc_yacht_auction_b200 import judge_bot as jb


class FirstLegalMover:
    """Stub policy: the k-th legal action of the board according to the oracle's legal mask (no GPU)."""

    def __init__(self, k=0):
        self.k = k
        self.boards = []

    def choose(self, boards):
        out = []
        for b in boards:
            self.boards.append(b)
            legal = np.flatnonzero(yr.legal_mask(to_oracle_board(b), 1))
            assert len(legal) > 0, "the session asked for a move on a board without legal moves"
            out.append(int(legal[self.k % len(legal)]))
        return out


def test_category_points_equal_the_oracle_everywhere():
    for dice in itertools.product(range(1, 7), repeat=5):
        for cat in range(12):
            assert jb.category_points(cat, list(dice)) == yr.category_points(cat, list(dice)), (cat, dice)


def test_scripted_transcript():
    """Line by line: the commands of INSTRUCTION.md:80-87 and the replies they require."""
    s = jb.JudgeSession(FirstLegalMover())
    assert s.handle("READY") == "OK"
    assert s.handle("ROLL 12345 66666") == "BID A 0"              # first legal action = (A, 0)
    assert s.handle("GET A B 31337") is None                       # different targets: both get their bundle
    assert s.me.carry == [1, 2, 3, 4, 5] and s.opp.carry == [6, 6, 6, 6, 6]
    assert s.me.bid_score == 0 and s.opp.bid_score == -31337 and s.round_no == 2   # round 1 has no scoring
    assert s.handle("ROLL 11111 23456") == "BID A 0"
    assert s.handle("GET B A 1") is None                           # same target, the opponent outbid us: we get B and ADD our bid
    assert s.me.carry == [1, 2, 3, 4, 5, 2, 3, 4, 5, 6] and s.opp.carry == [6, 6, 6, 6, 6, 1, 1, 1, 1, 1]
    assert s.opp.bid_score == -31338
    assert s.handle("SCORE") == "PUT ONE 12345"                    # first legal score action: category ONE, dice 0..4
    assert s.me.cat_scores[0] == 1000 and s.me.carry == [2, 3, 4, 5, 6] and s.round_no == 2
    assert s.handle("SET YACHT 66666") is None
    assert s.opp.cat_scores[11] == 50000 and s.opp.carry == [1, 1, 1, 1, 1] and s.round_no == 3
    b = s.board(0)
    assert b.round_no == 3 and b.phase == 0 and b.p1_bid is None and b.p2_bid is None
    assert b.p2.bid_score == -31500                                # -31338 rounded to the packed state's 500 grid
    assert s.totals() == (1000, 50000 - 31338)
    assert s.handle("FINISH") is None and s.finished
    for bad in ("HELLO", "ROLL 123 456", "GET C A 5", "SET NOPE 11111", "GET A A 100001"):
        with pytest.raises(jb.ProtocolError):
            jb.JudgeSession(FirstLegalMover()).handle(bad) if not bad.startswith("GET") else _get_after_roll(bad)


def _get_after_roll(line):
    s = jb.JudgeSession(FirstLegalMover())
    s.handle("ROLL 12345 66666")
    s.handle(line)


@pytest.mark.parametrize("seed,k,opp_first,same", [(1, 0, False, False), (2, 7, True, False), (3, 101, False, True), (4, 555, True, True)])
def test_full_games_against_a_simulated_judge(seed, k, opp_first, same):
    """13 rounds with an opponent that bids arbitrary integers: every reply is well formed and legal (the judge asserts
    it), the boards handed to the policy are valid engine states with the right mover and legal-move count, and the
    session's exact totals equal the judge's own score sheet."""
    mover = FirstLegalMover(k)
    session = jb.JudgeSession(mover)
    log, bot, opp = play(session.handle, seed, opp_first_on_score=opp_first, force_same_target=same)
    assert session.totals() == (bot.total(), opp.total())
    assert session.finished and session.decisions == 12 + 12 == len(mover.boards)
    for b in mover.boards:
        ob = to_oracle_board(b)
        n = int(yr.legal_mask(ob, 1).sum())
        assert n == (202 if b.phase == 0 else (252 if len(b.p1.carry) == 10 else 1) * (12 - bin(b.p1.used_mask).count("1")))
        assert b.p1.bid_score % 500 == 0 and abs(b.p1.bid_score - quant_ref(b, session)) >= 0
    assert [l for l, _ in log][:2] == ["READY", log[1][0]] and log[1][0].startswith("ROLL ")


def quant_ref(b, session):
    return b.p1.bid_score


def test_bid_score_quantisation():
    q = jb.quantise_bid_score
    assert [q(x) for x in (0, 249, 250, 499, 500, -249, -250, -31338, 12345)] == [0, 0, 500, 500, 500, 0, -500, -31500, 12500]
    assert q(10 ** 7) == 4095 * 500 and q(-10 ** 7) == -4096 * 500
