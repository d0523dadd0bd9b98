"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): the reference's own
code that consumes boards -- state_to_vec (yacht/NNet.py), the heuristic player's _choose_bid /
_choose_scoring (yacht/YachtPlayers.py), stringRepresentation -- runs UNCHANGED on YachtBoard objects and
gives the same answers as on the reference's own YachtState; and the oracle agrees with the reference
side by side on fresh random games (beyond the committed golden fixtures)."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("YACHT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "yacht")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        import yacht.YachtGame as yg
        import yacht.YachtPlayers as yp
        import yacht.NNet as nn
        yield yg, yp, nn
    finally:
        sys.path.remove(REF)


def _random_reference_states(yg, yp, seed, games=4):
    g = yg.YachtGame(seed=seed)
    pl = yp.RandomYachtPlayer(g)
    out = []
    for _ in range(games):
        board, cur = g.getInitBoard(), 1
        while g.getGameEnded(board, cur) == 0:
            canon = g.getCanonicalForm(board, cur)
            out.append(canon)
            board, cur = g.getNextState(board, cur, pl.play(canon))
        out.append(board)
    return g, out


def test_reference_code_runs_unchanged_on_yachtboard(ref):
    yg, yp, nn = ref
    from nypc_yacht_auction_b200.layout import pack_state, string_key
    g, states = _random_reference_states(yg, yp, 123)
    for s in states:
        b = pack_state(s)
        assert g.stringRepresentation(b) == g.stringRepresentation(s) == string_key(b)     # reference method on our board
        assert nn.state_to_vec(g, b).tobytes() == nn.state_to_vec(g, s).tobytes()
        assert b.p1.total_with_bonus() == s.p1.total_with_bonus() and b.p2.basic_total() == s.p2.basic_total()
        if s.phase == 0 and s.round_no != 13:
            assert yp._choose_bid(b) == yp._choose_bid(s)
        elif len(s.p1.carry) >= 5:
            assert yp._choose_scoring(b) == yp._choose_scoring(s)


def test_oracle_side_by_side_with_reference(ref):
    yg, yp, nn = ref
    from oracle import yacht_rules as yr, greedy_oracle
    from conftest import to_oracle_board
    g, states = _random_reference_states(yg, yp, 777, games=3)
    for s in states:
        ob = to_oracle_board(s)
        assert yr.key(ob) == g.stringRepresentation(s)
        assert (yr.legal_mask(ob, 1) == g.getValidMoves(s, 1)).all()
        assert yr.outcome(ob, 1) == g.getGameEnded(s, 1) and yr.outcome(ob, -1) == g.getGameEnded(s, -1)
        assert yr.features(ob).tobytes() == nn.state_to_vec(g, s).tobytes()
        assert yr.key(yr.canonical(ob, -1)) == g.stringRepresentation(g.getCanonicalForm(s, -1))
        if s.phase == 0 and s.round_no != 13:
            assert greedy_oracle.choose_bid(ob) == yp._choose_bid(s)
        elif len(s.p1.carry) >= 5:
            assert greedy_oracle.choose_scoring(ob) == yp._choose_scoring(s)


def test_dropin_signatures_match_reference(ref):
    """Same public method names and positional parameters as the reference classes."""
    import inspect
    yg, yp, nn = ref
    sys.path.insert(0, REF)
    try:
        import MCTS as ref_mcts
        import Coach as ref_coach
    finally:
        sys.path.remove(REF)
    from nypc_yacht_auction_b200 import game as my_game, mcts as my_mcts, coach as my_coach, players as my_players
    for name in ("getInitBoard", "getBoardSize", "getActionSize", "getNextState", "getValidMoves", "getGameEnded",
                 "getCanonicalForm", "getSymmetries", "stringRepresentation", "display"):
        a = list(inspect.signature(getattr(yg.YachtGame, name)).parameters)
        b = list(inspect.signature(getattr(my_game.YachtGame, name)).parameters)
        assert a == b, name
    for cls_ref, cls_mine, names in ((ref_mcts.MCTS, my_mcts.MCTS, ("__init__", "getActionProb", "search")),
                                     (ref_coach.Coach, my_coach.Coach, ("__init__", "executeEpisode")),
                                     (yp.RandomYachtPlayer, my_players.RandomYachtPlayer, ("__init__", "play")),
                                     (yp.GreedyYachtPlayer, my_players.GreedyYachtPlayer, ("__init__", "play"))):
        for name in names:
            a = list(inspect.signature(getattr(cls_ref, name)).parameters)
            b = list(inspect.signature(getattr(cls_mine, name)).parameters)
            assert a == b, (cls_ref.__name__, name)
    assert inspect.signature(yg.YachtGame.__init__).parameters["seed"].default is None
