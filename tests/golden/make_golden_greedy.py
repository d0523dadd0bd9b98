#!/usr/bin/env python
"""Golden vectors for the heuristic player from the UNMODIFIED reference
(yacht/YachtPlayers.py, imported from /root/reference; build container only).
Re-run with:  python tests/golden/make_golden_greedy.py"""
import hashlib
import json
import os
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

from yacht.YachtGame import YachtGame                                   # noqa: E402
from yacht.YachtPlayers import GreedyYachtPlayer, RandomYachtPlayer, _choose_bid, _choose_scoring   # noqa: E402


def play(seed, p1_kind, p2_kind):
    """Arena.playGame's loop (Arena.py:49-71) with the reference players; records what greedy chose."""
    g = YachtGame(seed=seed)
    mk = lambda k: GreedyYachtPlayer(g) if k == "greedy" else RandomYachtPlayer(g)
    players = {1: mk(p1_kind), -1: mk(p2_kind)}
    kinds = {1: p1_kind, -1: p2_kind}
    board = g.getInitBoard()
    cur = 1
    plies = []
    h = hashlib.sha256()
    while g.getGameEnded(board, cur) == 0:
        canon = g.getCanonicalForm(board, cur)
        raw = None
        if kinds[cur] == "greedy":
            raw = int(_choose_bid(canon)) if (canon.phase == 0 and canon.round_no != 13) else int(_choose_scoring(canon))
        a = int(players[cur].play(canon))
        valids = g.getValidMoves(canon, 1)
        assert valids[a] > 0
        h.update(g.stringRepresentation(canon).encode())
        h.update(str(a).encode())
        plies.append({"key": g.stringRepresentation(canon), "kind": kinds[cur], "raw": raw, "action": a})
        board, cur = g.getNextState(board, cur, a)
    return {"seed": seed, "p1": p1_kind, "p2": p2_kind, "plies": plies, "final_key": g.stringRepresentation(board),
            "totals": [int(board.p1.total_with_bonus()), int(board.p2.total_with_bonus())], "sha256": h.hexdigest(),
            "rng_after": int(np.random.randint(0, 2 ** 31))}


def main():
    games = [play(0, "greedy", "greedy"), play(1, "greedy", "greedy"), play(2, "greedy", "random"),
             play(3, "random", "greedy"), play(4, "greedy", "greedy"), play(5, "greedy", "random")]
    with open(os.path.join(HERE, "greedy_golden.json"), "w") as f:
        json.dump({"numpy": np.__version__, "games": games}, f, separators=(",", ":"))
    for gm in games:
        raws = [p for p in gm["plies"] if p["kind"] == "greedy"]
        fb = sum(1 for p in raws if p["raw"] != p["action"])
        print(gm["seed"], gm["p1"], gm["p2"], gm["totals"], "greedy plies", len(raws), "fallbacks", fb,
              "max raw bid action", max(p["raw"] for p in raws if p["raw"] < 400), gm["sha256"][:12])


if __name__ == "__main__":
    main()
