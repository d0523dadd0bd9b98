"""Pins the bulk C oracle (oracle/yacht_oracle.c) against the golden vectors made from the reference and
against the pure-Python oracle; also checks its independently written pack routine against the
product's documented packed layout."""
import hashlib
import itertools

import numpy as np

from oracle import c_oracle, philox
from oracle import yacht_rules as yr
from test_oracle_vs_golden import play_philox_game


def test_philox_and_tables(rules_golden):
    meta, arr = rules_golden
    assert [hex(x) for x in c_oracle.philox((0, 0, 0, 0), (0, 0))] == meta["philox_kat"]["zero"]
    assert [hex(x) for x in c_oracle.philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0))] == \
        meta["philox_kat"]["pi"]
    table = np.asarray([[c_oracle.category_points(c, list(d)) for c in range(12)]
                        for d in itertools.product(range(1, 7), repeat=5)], dtype=np.int32)
    assert hashlib.sha256(table.tobytes()).hexdigest() == meta["score_table_sha256_int32"]
    lib = c_oracle.load()
    subs = np.asarray([[lib.yo_subset_position(s, i) for i in range(5)] for s in range(252)], dtype=np.uint8)
    assert hashlib.sha256(subs.tobytes()).hexdigest() == meta["comb_sha256_uint8"]


def test_reference_driven_games(rules_golden):
    """The reference itself under injected Philox dice (golden) vs the C oracle's batched driver."""
    meta, _ = rules_golden
    for tr in meta["philox"]:
        r = c_oracle.play_random(1, 48, tr["seed"], tr["game"], want_keys=True)
        assert r["actions"][:, 0].tolist() == tr["actions"]
        assert r["keys"][0] == tr["final_key"]
        assert int(r["result"][47, 0]) == int(tr["ended_p1"])
        assert r["steps"] == 48


def test_matches_python_oracle_and_documented_layout():
    from nypc_yacht_auction_b200.layout import pack_state
    from test_layout import parse_key
    seed, base, n = 31, 1000, 12
    r = c_oracle.play_random(n, 48, seed, base, want_keys=True)
    for g in range(n):
        board, actions, _ = play_philox_game(seed, base + g)
        assert r["actions"][:, g].tolist() == actions
        assert r["keys"][g] == yr.key(board)
        assert tuple(int(x) for x in r["packed"][47, g]) == pack_state(parse_key(yr.key(board))).words
    # legal-count schedule
    expect = [202, 202] + sum(([202, 202, (14 - rr) * 252, (14 - rr) * 252] for rr in range(2, 13)), []) + [1, 1]
    assert (r["legal"] == np.asarray(expect)[:, None]).all()


def test_auto_reset_and_bulk():
    r = c_oracle.play_random(2000, 100, 5, 0, auto_reset=True)
    assert r["steps"] == 2000 * 100
    assert (r["result"][47] != 0).all() and (r["result"][95] != 0).all() and (r["result"][:47] == 0).all()
    assert (r["packed"][48, :, 0] & 15 == 1).all()          # ply 48 is the first ply of the next episode (round 1)
