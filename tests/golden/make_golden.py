#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (imported from
/root/reference, which exists only in the build container) and store them as small
fixtures next to this script.  Re-run with:

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own; these outputs of the
reference are what pins ``oracle/`` (and through it the CUDA path).

Outputs
  rules_golden.npz   score table 7776x12 (u8, /1000), subset table 252x5, feature rows
  rules_golden.json  sha256 digests, seeded MT19937 traces (actions, recorded dice, per-ply
                     keys, bit-packed legal masks, outcomes), Philox-injected traces
"""
import hashlib
import itertools
import json
import os
import random
import sys

import numpy as np

# The fixtures pin numpy >= 2 semantics (NEP 50: a Python float times an np.float32 stays float32), which decide the
# bits of the reference's UCB / Q arithmetic (MCTS.py:125-129,155-156); under numpy 1.x the reference itself behaves differently.
assert int(np.__version__.split(".")[0]) >= 2, "regenerate the golden files with numpy >= 2 (NEP 50), got %s" % np.__version__

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("YACHT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import yacht.YachtGame as ref_mod                      # noqa: E402
from yacht.YachtGame import YachtGame, score_category, COMB_5_OF_10   # noqa: E402
from yacht.YachtPlayers import RandomYachtPlayer       # noqa: E402
from yacht.NNet import state_to_vec                    # noqa: E402

from oracle import philox                               # noqa: E402  (protocol only: which dice to inject)


def score_table():
    rows = []
    for dice in itertools.product(range(1, 7), repeat=5):
        rows.append([score_category(c, list(dice)) for c in range(12)])
    return np.asarray(rows, dtype=np.int32)


class Recorder:
    """Wraps the reference's two RNG hooks, recording what they return (stream unchanged)."""

    def __init__(self):
        self.orig_roll = ref_mod.roll_five
        self.orig_tie = ref_mod.tiebreak_uniform
        self.log = []

    def __enter__(self):
        def roll():
            r = [int(x) for x in self.orig_roll()]
            self.log.append(["roll", r])
            return r

        def tie():
            t = int(self.orig_tie())
            self.log.append(["tie", t])
            return t
        ref_mod.roll_five = roll
        ref_mod.tiebreak_uniform = tie
        return self

    def __exit__(self, *a):
        ref_mod.roll_five = self.orig_roll
        ref_mod.tiebreak_uniform = self.orig_tie

    def take(self):
        out, self.log = self.log, []
        return out


def pack_mask(v):
    return np.packbits(np.asarray(v, dtype=np.uint8), bitorder="little").tobytes().hex()


def seeded_trace(seed):
    """YachtGame(seed) + RandomYachtPlayer on both sides, following Arena.playGame's call order
    (Arena.py:49-71) so the MT19937 stream is the one SURVEY.md section 8c quotes."""
    with Recorder() as rec:
        g = YachtGame(seed=seed)
        pl = RandomYachtPlayer(g)
        board = g.getInitBoard()
        init_draws = rec.take()
        cur = 1
        plies = []
        h = hashlib.sha256()
        feats = []
        while g.getGameEnded(board, cur) == 0:
            canon = g.getCanonicalForm(board, cur)
            a = pl.play(canon)
            valid = g.getValidMoves(canon, 1)
            ckey = g.stringRepresentation(canon)
            h.update(ckey.encode())
            h.update(valid.tobytes())
            feats.append(state_to_vec(g, canon))
            board, nxt = g.getNextState(board, cur, a)
            plies.append({
                "player": cur, "action": int(a), "canon_key": ckey, "mask": pack_mask(valid),
                "draws": rec.take(), "next_key": g.stringRepresentation(board), "next_player": int(nxt),
                "ended": float(g.getGameEnded(board, nxt)),
            })
            cur = nxt
    return {
        "seed": seed, "init_draws": init_draws, "plies": plies,
        "final_key": g.stringRepresentation(board),
        "totals": [board.p1.total_with_bonus(), board.p2.total_with_bonus()],
        "ended_p1": float(g.getGameEnded(board, 1)), "sha256": h.hexdigest(),
    }, np.asarray(feats, dtype=np.float32)


class PhiloxInjector:
    """Feeds the engine's draw protocol (oracle/philox.py docstring) into the reference."""

    def __init__(self, seed, game_id, episode=0):
        self.seed, self.game, self.episode = seed, game_id, episode
        self.cur = None
        self.used_a = False

    def arm(self, ply, tag, depth=0, sim=0):
        self.cur = philox.Draw(self.seed, self.game, self.episode, ply, tag, depth, sim)
        self.used_a = False

    def roll(self):
        if not self.used_a:
            self.used_a = True
            return self.cur.roll_a()
        return self.cur.roll_b()

    def tie(self):
        return self.cur.tie()


def philox_trace(seed, game_id):
    """Reference driven by injected Philox dice and the Philox random-legal policy."""
    inj = PhiloxInjector(seed, game_id)
    keep = (ref_mod.roll_five, ref_mod.tiebreak_uniform)
    ref_mod.roll_five, ref_mod.tiebreak_uniform = inj.roll, inj.tie
    try:
        g = YachtGame()
        inj.arm(0, philox.TAG_INIT)
        board = g.getInitBoard()
        cur, ply = 1, 0
        h = hashlib.sha256()
        actions = []
        while g.getGameEnded(board, cur) == 0:
            canon = g.getCanonicalForm(board, cur)
            valid = g.getValidMoves(canon, 1)
            legal = np.nonzero(valid)[0]
            pick = philox.Draw(seed, game_id, 0, ply, philox.TAG_ACTION).pick(len(legal))
            a = int(legal[pick])
            h.update(g.stringRepresentation(canon).encode())
            h.update(valid.tobytes())
            inj.arm(ply, philox.TAG_REAL)
            board, cur = g.getNextState(board, cur, a)
            actions.append(a)
            ply += 1
        return {
            "seed": seed, "game": game_id, "actions": actions, "final_key": g.stringRepresentation(board),
            "totals": [board.p1.total_with_bonus(), board.p2.total_with_bonus()],
            "ended_p1": float(g.getGameEnded(board, 1)), "sha256": h.hexdigest(),
        }
    finally:
        ref_mod.roll_five, ref_mod.tiebreak_uniform = keep


def main():
    table = score_table()
    comb = np.asarray(COMB_5_OF_10, dtype=np.uint8)
    out = {
        "numpy": np.__version__,
        "score_table_sha256_int32": hashlib.sha256(table.tobytes()).hexdigest(),
        "score_table_sum": int(table.sum()),
        "score_table_col_sums": [int(x) for x in table.sum(0)],
        "comb_sha256_uint8": hashlib.sha256(comb.tobytes()).hexdigest(),
        "philox_kat": {
            "zero": [hex(x) for x in philox.philox4x32_10((0, 0, 0, 0), (0, 0))],
            "ones": [hex(x) for x in philox.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2)],
            "pi": [hex(x) for x in philox.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                                                         (0xA4093822, 0x299F31D0))],
        },
        "seeded": [], "philox": [],
    }
    feats = {}
    for seed in (0, 1, 2, 3):
        tr, f = seeded_trace(seed)
        out["seeded"].append(tr)
        feats["features_seed%d" % seed] = f
    for seed, gid in ((0, 0), (0, 1), (0, 65535), (7, 3), (7, 123456789), (2 ** 40 + 5, 17)):
        out["philox"].append(philox_trace(seed, gid))
    assert (table % 1000 == 0).all() and table.max() == 50000
    np.savez_compressed(os.path.join(HERE, "rules_golden.npz"),
                        score_table_k=(table // 1000).astype(np.uint8), subsets=comb, **feats)
    with open(os.path.join(HERE, "rules_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("score sha", out["score_table_sha256_int32"], "sum", out["score_table_sum"])
    print("comb sha", out["comb_sha256_uint8"])
    print("philox kat", out["philox_kat"])
    for tr in out["seeded"]:
        print("seed", tr["seed"], len(tr["plies"]), "plies", tr["totals"], tr["ended_p1"], tr["sha256"][:16])
    for tr in out["philox"]:
        print("philox", tr["seed"], tr["game"], tr["totals"], tr["ended_p1"], tr["sha256"][:16])


if __name__ == "__main__":
    main()
