"""The drop-in a0g surface (YachtGame / MCTS / Coach.executeEpisode) on the GPU against golden runs of
the reference under the same MT19937 seeds: same hooks, same RNG consumption, same boards."""
import hashlib
import json
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import mcts_oracle
from conftest import to_oracle_board

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def arena_play_game(game, player1, player2, on_ply=None):
    """Restates the loop of Arena.playGame (Arena.py:30-93) to drive the Game API the way Arena does."""
    players = [player2, None, player1]
    cur = 1
    board = game.getInitBoard()
    while game.getGameEnded(board, cur) == 0:
        canon = game.getCanonicalForm(board, cur)
        action = players[cur + 1](canon)
        valids = game.getValidMoves(game.getCanonicalForm(board, cur), 1)
        assert valids[action] > 0
        if on_ply:
            on_ply(canon, valids, action)
        board, cur = game.getNextState(board, cur, action)
    return cur * game.getGameEnded(board, cur), board


def test_seeded_arena_games_match_reference(rules_golden):
    """YachtGame(seed) + RandomYachtPlayer: the MT19937-seeded traces quoted in SURVEY.md section 8c."""
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.players import RandomYachtPlayer
    meta, _ = rules_golden
    for tr in meta["seeded"]:
        g = YachtGame(seed=tr["seed"])
        pl = RandomYachtPlayer(g)
        h = hashlib.sha256()
        actions = []

        def on_ply(canon, valids, action):
            h.update(g.stringRepresentation(canon).encode())
            h.update(valids.tobytes())
            actions.append(action)
        # Arena evaluates the player before the legality check; the digest needs the same order
        result, board = arena_play_game(g, pl.play, pl.play, on_ply)
        assert actions == [p["action"] for p in tr["plies"]]
        assert h.hexdigest() == tr["sha256"]
        assert g.stringRepresentation(board) == tr["final_key"]
        assert [board.p1.total_with_bonus(), board.p2.total_with_bonus()] == tr["totals"]
        assert g.getGameEnded(board, 1) == tr["ended_p1"]
    assert meta["seeded"][0]["sha256"] == "1110e34fdc95c665983452581b2e4d862e3442376f14bd86e701e9326bf5a70c"


def test_game_api_contract():
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.layout import YachtBoard
    g = YachtGame(seed=3)
    assert g.getBoardSize() == (1, 59) and g.getActionSize() == 3226
    b = g.getInitBoard()
    assert isinstance(b, YachtBoard) and b.round_no == 1 and b.phase == 0 and len(b.rollA) == 5
    assert g.getCanonicalForm(b, 1) is b
    v = g.getValidMoves(b, 1)
    assert v.dtype == np.uint8 and v.shape == (3226,) and int(v.sum()) == 202
    assert g.getGameEnded(b, 1) == 0.0 and isinstance(g.getGameEnded(b, 1), float)
    assert g.getSymmetries(b, [1]) == [(b, [1])]
    nb, nxt = g.getNextState(b, 1, np.int64(7))
    assert nxt == -1 and nb.p1_bid == ("A", 3500) and b.p1_bid is None          # input untouched
    c = g.getCanonicalForm(nb, -1)
    assert c.p2_bid == ("A", 3500) and c.p1_bid is None
    with pytest.raises(ValueError):
        g.getNextState(b, 1, 202)
    with pytest.raises(ValueError):
        g.getNextState(b, 1, -1)
    assert pickle.loads(pickle.dumps(nb)) == nb
    assert g.stringRepresentation(nb).startswith("r1|ph0|A")
    vec = g.stateToVec(c)
    assert vec.dtype == np.float32 and vec.shape == (59,)
    from oracle import yacht_rules as yr
    assert vec.tobytes() == yr.features(to_oracle_board(c)).tobytes()


def test_coach_execute_episode_matches_reference():
    """Coach.executeEpisode of the reference (golden) vs the drop-in Coach + MCTS + YachtGame."""
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.coach import Coach
    with open(os.path.join(GOLDEN, "mcts_golden.json")) as f:
        cases = json.load(f)["coach"]

    class Args(dict):
        __getattr__ = dict.__getitem__

    class HashedNet:
        def predict(self, board):
            return mcts_oracle.hashed_evaluator(to_oracle_board(board), 5)

    for case in cases:
        g = YachtGame(seed=case["mt_seed"])
        args = Args(numMCTSSims=case["sims"], cpuct=case["cpuct"], tempThreshold=case["temp_threshold"],
                    search_seed=case["search_seed"], tree_id=case["tree_id"], search_dice="philox")
        coach = Coach(g, HashedNet(), args)
        examples = coach.executeEpisode()
        assert len(examples) == len(case["examples"])
        for (board, pi, v), ref in zip(examples, case["examples"]):
            assert g.stringRepresentation(board) == ref["key"]
            assert float(v) == ref["v"]
            assert {str(a): float(p).hex() for a, p in enumerate(pi) if p} == ref["pi"]
        assert int(np.random.randint(0, 2 ** 31)) == case["rng_after"]        # same global RNG consumption
        blob = pickle.dumps(examples)                                             # Coach.saveTrainExamples pickles these
        assert len(pickle.loads(blob)) == len(examples)


def test_batched_self_play_examples():
    """BatchedSelfPlay: labels follow Coach.py:69-72, policies are the root visit counts."""
    import torch
    from nypc_yacht_auction_b200.coach import BatchedSelfPlay
    sp = BatchedSelfPlay(5, 10, cpuct=1.5, seed=4, game_base=50, temp_threshold=15)
    out = sp.execute_episodes()
    assert out["features"].shape == (48, 5, 59) and out["value"].shape == (48, 5)
    res = out["result_p1"].cpu().numpy()
    val = out["value"].cpu().numpy()
    players = sp.ex_players.cpu().numpy()
    assert (val == res[None, :] * players).all()
    pi = BatchedSelfPlay.dense_policy(out["actions"], out["counts"])
    assert torch.allclose(pi.sum(-1), torch.ones_like(pi.sum(-1)))
    # oracle self-play of game 52 with the same uniform evaluator gives the same visit counts at ply 0..5
    trace = mcts_oracle.self_play_game(mcts_oracle.uniform_evaluator, 10, 1.5, 4, 52, max_plies=6)[0]
    for t in range(6):
        c = {int(a): int(n) for a, n in zip(out["actions"][t, 2].cpu().numpy(), out["counts"][t, 2].cpu().numpy()) if n}
        assert c == trace[t]["counts"]


def _wave_examples(total, wave, sims, evaluator, use_graph, seed, first):
    from nypc_yacht_auction_b200.coach import self_play_in_waves
    got = {}

    def on_wave(w, ex):
        got[w] = {k: v.clone() for k, v in ex.items()}
    totals = self_play_in_waves(total, wave, sims, evaluator, first_game=first, on_wave=on_wave, use_graph=use_graph, seed=seed)
    return got, totals


def _assert_waves_equal(got, ref, wave):
    for w in sorted(got):
        sl = slice(wave * w, wave * w + got[w]["result_p1"].shape[0])
        assert torch.equal(got[w]["result_p1"], ref["result_p1"][sl]), w
        assert torch.equal(got[w]["counts"], ref["counts"][:, sl]) and torch.equal(got[w]["actions"], ref["actions"][:, sl]), w
        assert torch.equal(got[w]["features"], ref["features"][:, sl]), w


def test_waves_equal_one_big_batch():
    """configs[4] driver: playing 3 waves of 4 games on one tree pool and a ragged last wave of 2 gives exactly the games of
    one batch of 14 (global game ids key the Philox streams), so sharding over waves / GPUs never changes results."""
    from nypc_yacht_auction_b200.coach import BatchedSelfPlay
    big = BatchedSelfPlay(14, 6, seed=9, game_base=200)
    ref = big.execute_episodes()
    got, totals = _wave_examples(14, 4, 6, None, True, 9, 200)
    assert sorted(got) == [0, 1, 2, 3] and got[3]["result_p1"].shape[0] == 2
    _assert_waves_equal(got, ref, 4)
    r = ref["result_p1"]
    assert totals == (int((r > 0.5).sum()), int((r < -0.5).sum()), int((r.abs() < 0.5).sum()))


def test_graphed_network_waves_equal_one_big_batch(monkeypatch):
    """The configs[4] path proper: network evaluator + the simulation wave replayed as a CUDA graph.  A captured
    launch freezes its by-value arguments, so the global game ids of the in-search Philox draws (dice rolled below
    round-1 roots, tie-breaks of equal bids; quirk Q1) must come from device memory -- otherwise every wave after the
    first searches with wave 0's ids (round-1 VERDICT weak #1).  Waves of 8 games x 24 sims equal one batch of 24
    games bit for bit, graphed or not; and as a negative control the same run with the device-side base left stale
    must DIFFER, which proves the comparison sees in-search draws."""
    from nypc_yacht_auction_b200.coach import BatchedSelfPlay
    from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    torch.manual_seed(11)
    net = YachtPolicyValueNet().cuda().eval()
    total, wave, sims, seed, first = 24, 8, 24, 13, 4000
    big = BatchedSelfPlay(total, sims, evaluator=FusedYachtEvaluator(net, total), seed=seed, game_base=first)
    ref = big.execute_episodes()                                         # eager, one batch: by-value game ids
    for use_graph in (True, False):
        got, totals = _wave_examples(total, wave, sims, FusedYachtEvaluator(net, wave), use_graph, seed, first)
        _assert_waves_equal(got, ref, wave)
    r = ref["result_p1"]
    assert totals == (int((r > 0.5).sum()), int((r < -0.5).sum()), int((r.abs() < 0.5).sum()))
    # negative control: freeze the device-side base at wave 0's value (what a by-value captured argument did)
    monkeypatch.setattr(BatchedMCTS, "sync_game_base", lambda self: None)
    stale, _ = _wave_examples(total, wave, sims, FusedYachtEvaluator(net, wave), True, seed, first)
    assert torch.equal(stale[0]["counts"], ref["counts"][:, :wave])      # wave 0 is unaffected
    assert not all(torch.equal(stale[w]["counts"], ref["counts"][:, wave * w:wave * w + wave]) for w in (1, 2))


def test_seeded_episode_matches_unpatched_reference():
    """Coach.executeEpisode of the UNPATCHED reference (only YachtGame(seed) set; in-search dice from the global
    numpy / random streams) vs the drop-in Coach + MCTS + YachtGame in their default mode: identical boards,
    policies, values and identical RNG states afterwards (both numpy's and random's)."""
    import random
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.coach import Coach
    with open(os.path.join(GOLDEN, "mcts_golden.json")) as f:
        cases = json.load(f)["coach_plain"]

    class Args(dict):
        __getattr__ = dict.__getitem__

    class HashedNet:
        def predict(self, board):
            return mcts_oracle.hashed_evaluator(to_oracle_board(board), 9)

    for case in cases:
        g = YachtGame(seed=case["mt_seed"])
        coach = Coach(g, HashedNet(), Args(numMCTSSims=case["sims"], cpuct=case["cpuct"], tempThreshold=case["temp_threshold"]))
        examples = coach.executeEpisode()
        assert len(examples) == len(case["examples"])
        for (board, pi, v), ref in zip(examples, case["examples"]):
            assert g.stringRepresentation(board) == ref["key"]
            assert float(v) == ref["v"]
            assert {str(a): float(p).hex() for a, p in enumerate(pi) if p} == ref["pi"]
        assert int(np.random.randint(0, 2 ** 31)) == case["rng_after"]
        assert random.random().hex() == case["py_rng_after"]


def test_examples_round_trip_through_reference_format(tmp_path):
    """device examples -> (board, pi, v) tuples (what Coach.executeEpisode returns and
    Coach.saveTrainExamples pickles, Coach.py:66-72,144-152) -> pickle -> tensors again."""
    from nypc_yacht_auction_b200 import examples as exm
    from nypc_yacht_auction_b200.coach import BatchedSelfPlay
    sp = BatchedSelfPlay(5, 6, evaluator=None, seed=9, record_states=True)
    ex = sp.execute_episodes()
    tuples = exm.to_reference_examples(ex)
    assert len(tuples) == 5 * 48
    board, pi, v = tuples[48 + 3]                                     # game 1, ply 3
    assert abs(sum(pi) - 1.0) < 1e-12 and len(pi) == 3226 and v in (1.0, -1.0, 1e-4, -1e-4)
    assert 1 <= board.round_no <= 13 and board.phase in (0, 1)
    path = tmp_path / "checkpoint_0.pth.tar.examples"
    exm.save_train_examples(str(path), [tuples])
    history = exm.load_train_examples(str(path))
    back = exm.from_reference_examples(history[0])
    feats = ex["features"].permute(1, 0, 2).reshape(-1, 59)           # game-major, like the tuple list
    assert torch.equal(back["features"], feats)
    dense = BatchedSelfPlay.dense_policy(ex["actions"], ex["counts"]).permute(1, 0, 2).reshape(-1, 3226)
    assert torch.allclose(back["pi"].double(), dense, atol=1e-7)
    assert torch.equal(back["value"], ex["value"].permute(1, 0).reshape(-1))
