"""Device greedy player (ya_greedy_action) against the reference's GreedyYachtPlayer (golden) and the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import greedy_oracle
from conftest import to_oracle_board
from test_layout import parse_key

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "greedy_golden.json")


def _games():
    with open(GOLDEN) as f:
        return json.load(f)["games"]


def test_kernel_choices_match_reference_on_every_golden_state():
    from nypc_yacht_auction_b200.engine import BatchedYacht
    from nypc_yacht_auction_b200.layout import pack_state
    plies = [p for gm in _games() for p in gm["plies"] if p["kind"] == "greedy"]
    boards = [pack_state(parse_key(p["key"])) for p in plies]
    env = BatchedYacht(len(boards))
    env.load_boards(boards, [1] * len(boards))
    raw = torch.zeros(len(boards), dtype=torch.int32, device="cuda")
    acts = env.greedy_actions(raw=raw, fallback=False).cpu().numpy()
    raw = raw.cpu().numpy()
    assert raw.tolist() == [p["raw"] for p in plies]
    assert acts.tolist() == [p["raw"] if p["raw"] == p["action"] else -1 for p in plies]


def test_dropin_greedy_player_reproduces_seeded_games():
    """Arena-style loop with the drop-in YachtGame + GreedyYachtPlayer / RandomYachtPlayer under the
    reference's MT19937 seeds: identical actions, boards and RNG consumption."""
    from nypc_yacht_auction_b200.game import YachtGame
    from nypc_yacht_auction_b200.players import GreedyYachtPlayer, RandomYachtPlayer
    for gm in _games()[:4]:
        g = YachtGame(seed=gm["seed"])
        mk = lambda k: GreedyYachtPlayer(g) if k == "greedy" else RandomYachtPlayer(g)
        players = {1: mk(gm["p1"]), -1: mk(gm["p2"])}
        board, cur = g.getInitBoard(), 1
        h = hashlib.sha256()
        for p in gm["plies"]:
            canon = g.getCanonicalForm(board, cur)
            assert g.stringRepresentation(canon) == p["key"]
            a = int(players[cur].play(canon))
            assert a == p["action"]
            valids = g.getValidMoves(canon, 1)
            assert valids[a] > 0
            h.update(g.stringRepresentation(canon).encode())
            h.update(str(a).encode())
            board, cur = g.getNextState(board, cur, a)
        assert h.hexdigest() == gm["sha256"]
        assert g.stringRepresentation(board) == gm["final_key"]
        assert int(np.random.randint(0, 2 ** 31)) == gm["rng_after"]


def test_batched_greedy_vs_oracle_with_overflow_quirk():
    """Random mid-game states + the bid-overflow state (quirk Q11): raw choice equals the oracle; with the
    device fallback the returned action is always legal."""
    from nypc_yacht_auction_b200.engine import BatchedYacht
    from nypc_yacht_auction_b200.layout import pack_state
    n = 400
    env = BatchedYacht(n, seed=8, game_base=10)
    masks = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    raw = torch.zeros(n, dtype=torch.int32, device="cuda")
    for ply in range(48):
        if ply % 5 == 0:
            acts = env.greedy_actions(raw=raw, fallback=True).cpu().numpy()
            env.valid_moves(out=masks)
            assert bool(masks.gather(1, torch.from_numpy(acts).long().cuda().unsqueeze(1)).all())
            boards = env.boards()
            pl = env.players.cpu().numpy()
            r = raw.cpu().numpy()
            from oracle import yacht_rules as yr
            for g in range(0, n, 7):
                ob = yr.canonical(to_oracle_board(boards[g]), int(pl[g]))
                assert greedy_oracle.greedy_action(ob)[0] == int(r[g]), (ply, g)
        env.play_ply(masks=None, auto_reset=False)
    q11 = pack_state(parse_key(
        "r5|ph0|A66666|B12345|p1b-|p2b-|p1c11234|p2c23456|p1u7|p2u7|p1s0,0,0,0,0,0,0,0,0,0,0,0|"
        "p2s5000,10000,15000,0,0,0,0,0,0,0,0,0|p1bid-400000|p2bid300000"))
    env1 = BatchedYacht(1)
    env1.load_boards([q11], [1])
    raw1 = torch.zeros(1, dtype=torch.int32, device="cuda")
    a = int(env1.greedy_actions(raw=raw1, fallback=True).item())
    assert int(raw1.item()) == greedy_oracle.greedy_action(to_oracle_board(q11))[0] and int(raw1.item()) >= 101
    assert 0 <= a < 202
