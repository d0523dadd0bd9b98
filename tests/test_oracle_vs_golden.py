"""Pins oracle/yacht_rules.py and oracle/philox.py against outputs of the reference itself
(tests/golden/rules_golden.*, produced by tests/golden/make_golden.py from /root/reference)."""
import hashlib
import itertools

import numpy as np

from oracle import philox
from oracle import yacht_rules as yr
from conftest import TapeDraw


def unpack_mask(hexstr):
    return np.unpackbits(np.frombuffer(bytes.fromhex(hexstr), dtype=np.uint8), bitorder="little")[:yr.N_ACTION]


def test_philox_known_answers(rules_golden):
    meta, _ = rules_golden
    # Random123 kat_vectors for philox4x32-10
    assert meta["philox_kat"]["zero"] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in philox.philox4x32_10((0, 0, 0, 0), (0, 0))] == meta["philox_kat"]["zero"]
    assert [hex(x) for x in philox.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2)] == meta["philox_kat"]["ones"]
    assert [hex(x) for x in philox.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                                                  (0xA4093822, 0x299F31D0))] == meta["philox_kat"]["pi"]


def test_score_table_exhaustive(rules_golden):
    meta, arr = rules_golden
    table = np.asarray([[yr.category_points(c, list(d)) for c in range(12)]
                        for d in itertools.product(range(1, 7), repeat=5)], dtype=np.int32)
    assert hashlib.sha256(table.tobytes()).hexdigest() == meta["score_table_sha256_int32"]
    assert (table // 1000 == arr["score_table_k"]).all()
    assert int(table.sum()) == 305745000


def test_subset_table(rules_golden):
    meta, arr = rules_golden
    mine = np.asarray(yr.SUBSETS, dtype=np.uint8)
    assert hashlib.sha256(mine.tobytes()).hexdigest() == meta["comb_sha256_uint8"]
    assert (mine == arr["subsets"]).all()


def test_seeded_traces_replay(rules_golden):
    meta, arr = rules_golden
    for tr in meta["seeded"]:
        board = yr.new_game(TapeDraw(tr["init_draws"]))
        cur = 1
        h = hashlib.sha256()
        feats = arr["features_seed%d" % tr["seed"]]
        for i, ply in enumerate(tr["plies"]):
            assert cur == ply["player"]
            assert yr.outcome(board, cur) == 0.0
            canon = yr.canonical(board, cur)
            assert yr.key(canon) == ply["canon_key"]
            mask = yr.legal_mask(canon, 1)
            assert (mask == unpack_mask(ply["mask"])).all()
            assert int(mask.sum()) == yr.legal_count(canon, 1)
            assert mask[ply["action"]] == 1
            f = yr.features(canon)
            assert f.dtype == np.float32 and f.tobytes() == feats[i].tobytes()
            h.update(yr.key(canon).encode())
            h.update(mask.tobytes())
            tape = TapeDraw(ply["draws"])
            board, cur = yr.next_state(board, cur, ply["action"], tape)
            assert tape.done()
            assert yr.key(board) == ply["next_key"]
            assert cur == ply["next_player"]
            assert yr.outcome(board, cur) == ply["ended"]
        assert h.hexdigest() == tr["sha256"]
        assert yr.key(board) == tr["final_key"]
        assert [s.total() for s in board.sides] == tr["totals"]
        assert yr.outcome(board, 1) == tr["ended_p1"]
    # the two digests SURVEY.md section 8c quotes
    assert meta["seeded"][0]["sha256"] == "1110e34fdc95c665983452581b2e4d862e3442376f14bd86e701e9326bf5a70c"
    assert meta["seeded"][1]["sha256"] == "4d2bacc83d68eb69a679da34eec4256a5aed8a08a4fc2f7a46f27fff19300e9e"


def play_philox_game(seed, gid, episode=0):
    board = yr.new_game(philox.Draw(seed, gid, episode, 0, philox.TAG_INIT))
    cur, ply = 1, 0
    h = hashlib.sha256()
    actions = []
    while yr.outcome(board, cur) == 0.0:
        canon = yr.canonical(board, cur)
        mask = yr.legal_mask(canon, 1)
        a = yr.random_legal_action(canon, 1, philox.Draw(seed, gid, episode, ply, philox.TAG_ACTION))
        assert mask[a] == 1
        h.update(yr.key(canon).encode())
        h.update(mask.tobytes())
        board, cur = yr.next_state(board, cur, a, philox.Draw(seed, gid, episode, ply, philox.TAG_REAL))
        actions.append(a)
        ply += 1
    return board, actions, h.hexdigest()


def test_philox_injected_traces(rules_golden):
    meta, _ = rules_golden
    for tr in meta["philox"]:
        board, actions, digest = play_philox_game(tr["seed"], tr["game"])
        assert actions == tr["actions"]
        assert digest == tr["sha256"]
        assert yr.key(board) == tr["final_key"]
        assert [s.total() for s in board.sides] == tr["totals"]
        assert yr.outcome(board, 1) == tr["ended_p1"]


def test_illegal_actions_match_reference_semantics():
    b = yr.new_game(philox.Draw(1, 2, 0, 0, philox.TAG_INIT))
    d = philox.Draw(1, 2, 0, 0, philox.TAG_REAL)
    for bad in (-1, 202, 3225):
        try:
            yr.next_state(b, 1, bad, d)
            assert False
        except ValueError:
            pass
    # score phase: used category / subset that does not fit -> silent no-op, turn passes
    b.phase = yr.SCORE
    b.rnd = 5
    b.sides[0].dice = [1, 2, 3, 4, 5]
    b.sides[0].used = 1
    nb, nxt = yr.next_state(b, 1, 202 + 0 * 252 + 0, d)
    assert nxt == -1 and yr.key(nb) == yr.key(b)
    nb, nxt = yr.next_state(b, 1, 202 + 3 * 252 + 1, d)      # subset (0,1,2,3,5) needs 6 dice
    assert nxt == -1 and yr.key(nb) == yr.key(b)
    try:
        yr.next_state(b, 1, 5, d)
        assert False
    except ValueError:
        pass
