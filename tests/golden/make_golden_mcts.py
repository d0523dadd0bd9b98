#!/usr/bin/env python
"""Golden MCTS traces from the UNMODIFIED reference (MCTS.py + yacht/YachtGame.py imported from
/root/reference; build container only).  Re-run with:  python tests/golden/make_golden_mcts.py

The reference draws dice inside search from the global numpy RNG; here its two RNG hooks
(yacht.YachtGame.roll_five / tiebreak_uniform) are replaced by the engine's Philox draw protocol
(oracle/philox.py), keyed by (ply of the root, sim index, search depth) which a wrapper around
MCTS.search tracks.  The evaluator is a deterministic pseudo-network (oracle/mcts_oracle.py).
The game loop follows Coach.executeEpisode (Coach.py:34-72) with the engine's count-sampling rule.
"""
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

import yacht.YachtGame as ref_mod                     # noqa: E402
from yacht.YachtGame import YachtGame                  # noqa: E402
from MCTS import MCTS                                   # noqa: E402
from utils import dotdict                               # noqa: E402

from oracle import philox, mcts_oracle                  # noqa: E402
from oracle import yacht_rules as yr                    # noqa: E402


def to_oracle_board(s):
    b = yr.Board()
    b.rnd, b.phase = s.round_no, s.phase
    b.pool_a, b.pool_b = [int(x) for x in s.rollA], [int(x) for x in s.rollB]
    conv = lambda x: None if x is None else ("AB".index(x[0]), x[1])
    b.bids = [conv(s.p1_bid), conv(s.p2_bid)]
    for i, p in enumerate((s.p1, s.p2)):
        b.sides[i] = yr.Side([int(x) for x in p.carry], p.used_mask, list(p.cat_scores), p.bid_score)
    return b


class Net:
    def __init__(self, fn):
        self.fn = fn
        self.calls = 0

    def predict(self, board):
        self.calls += 1
        return self.fn(to_oracle_board(board))


class Injector:
    def __init__(self, seed, game, episode=0):
        self.seed, self.game, self.episode = seed, game, episode
        self.cur, self.first = None, True

    def arm(self, ply, tag, depth=0, sim=0):
        self.cur = philox.Draw(self.seed, self.game, self.episode, ply, tag, depth, sim)
        self.first = True

    def roll(self):
        if self.first:
            self.first = False
            return self.cur.roll_a()
        return self.cur.roll_b()

    def tie(self):
        return self.cur.tie()


def qrepr(q):
    return ["f32" if isinstance(q, np.float32) else "f64", float(q).hex()]


def run_case(name, fn, sims, cpuct, seed, game, temp_threshold, max_plies=None):
    inj = Injector(seed, game)
    keep = (ref_mod.roll_five, ref_mod.tiebreak_uniform)
    ref_mod.roll_five, ref_mod.tiebreak_uniform = inj.roll, inj.tie
    try:
        g = YachtGame()
        net = Net(fn)
        mcts = MCTS(g, net, dotdict({"numMCTSSims": sims, "cpuct": cpuct}))
        state = {"ply": 0, "sim": -1, "depth": -1}
        orig = MCTS.search

        def search(board):
            state["depth"] += 1
            if state["depth"] == 0:
                state["sim"] += 1
            inj.arm(state["ply"], philox.TAG_SEARCH, state["depth"], state["sim"])
            try:
                return orig(mcts, board)
            finally:
                state["depth"] -= 1
        mcts.search = search

        inj.arm(0, philox.TAG_INIT)
        board = g.getInitBoard()
        cur, ply = 1, 0
        trace = []
        while True:
            canon = g.getCanonicalForm(board, cur)
            state.update(ply=ply, sim=-1, depth=-1)
            temp = int((ply + 1) < temp_threshold)
            mcts.getActionProb(canon, temp=1)        # temp only shapes the returned list; counts are what we record
            s = g.stringRepresentation(canon)
            counts = np.array([mcts.Nsa.get((s, a), 0) for a in range(g.getActionSize())], dtype=np.int64)
            word = philox.draw_words(seed, game, 0, ply, philox.TAG_ACTION)[3]
            if temp == 0:
                best = np.flatnonzero(counts == counts.max())
                action = int(best[(word * len(best)) >> 32])
            else:
                action = mcts_oracle.sample_from_counts(counts, word)
            trace.append({
                "ply": ply, "player": cur, "key": s, "action": action, "nodes": len(mcts.Ps), "ns": int(mcts.Ns[s]),
                "counts": {str(int(a)): int(counts[a]) for a in np.flatnonzero(counts)},
                "q": {str(int(a)): qrepr(mcts.Qsa[(s, int(a))]) for a in np.flatnonzero(counts)},
            })
            inj.arm(ply, philox.TAG_REAL)
            board, cur = g.getNextState(board, cur, action)
            ply += 1
            r = g.getGameEnded(board, cur)
            if r != 0 or (max_plies is not None and ply >= max_plies):
                break
        return {"name": name, "sims": sims, "cpuct": cpuct, "seed": seed, "game": game, "temp_threshold": temp_threshold,
                "max_plies": max_plies, "trace": trace, "final_key": g.stringRepresentation(board), "result": float(r),
                "final_player": cur, "leaf_evals": net.calls, "total_nodes": len(mcts.Ps), "terminal_states": len(mcts.Es)}
    finally:
        ref_mod.roll_five, ref_mod.tiebreak_uniform = keep


def coach_episode_case(mt_seed, sims, cpuct, temp_threshold, search_seed, tree_id):
    """Reference Coach.executeEpisode (Coach.py:34-72), unmodified, MT19937-seeded; only the dice rolled
    INSIDE MCTS.search are replaced by the Philox protocol (real moves keep the reference's own hooks)."""
    from Coach import Coach
    keep = (ref_mod.roll_five, ref_mod.tiebreak_uniform)
    inj = Injector(search_seed, tree_id)
    state = {"ply": -1, "sim": -1, "depth": -1}

    def roll():
        return inj.roll() if state["depth"] >= 0 else keep[0]()

    def tie():
        return inj.tie() if state["depth"] >= 0 else keep[1]()
    ref_mod.roll_five, ref_mod.tiebreak_uniform = roll, tie
    try:
        class HashedNet:
            def __init__(self, game=None, args=None):
                pass

            def predict(self, board):
                return mcts_oracle.hashed_evaluator(to_oracle_board(board), 5)

        g = YachtGame(seed=mt_seed)
        args = dotdict({"numMCTSSims": sims, "cpuct": cpuct, "tempThreshold": temp_threshold})
        coach = Coach(g, HashedNet(), args)
        mcts = coach.mcts
        orig_search, orig_gap = MCTS.search, MCTS.getActionProb

        def search(board):
            state["depth"] += 1
            if state["depth"] == 0:
                state["sim"] += 1
            inj.arm(state["ply"], philox.TAG_SEARCH, state["depth"], state["sim"])
            try:
                return orig_search(mcts, board)
            finally:
                state["depth"] -= 1

        def gap(board, temp=1):
            state["ply"] += 1
            state["sim"] = -1
            return orig_gap(mcts, board, temp=temp)
        mcts.search, mcts.getActionProb = search, gap
        examples = coach.executeEpisode()
        out = []
        for board, pi, v in examples:
            out.append({"key": g.stringRepresentation(board), "v": float(v),
                        "pi": {str(a): float(p).hex() for a, p in enumerate(pi) if p}})
        return {"mt_seed": mt_seed, "sims": sims, "cpuct": cpuct, "temp_threshold": temp_threshold,
                "search_seed": search_seed, "tree_id": tree_id, "examples": out,
                "rng_after": int(np.random.randint(0, 2 ** 31))}
    finally:
        ref_mod.roll_five, ref_mod.tiebreak_uniform = keep


def coach_plain_case(mt_seed, sims, cpuct, temp_threshold):
    """The reference exactly as shipped: Coach.executeEpisode seeded through YachtGame(seed); every dice roll
    and tie-break (real moves AND inside MCTS.search) comes from the global numpy / random streams."""
    import random
    from Coach import Coach

    class HashedNet:
        def __init__(self, game=None, args=None):
            pass

        def predict(self, board):
            return mcts_oracle.hashed_evaluator(to_oracle_board(board), 9)

    g = YachtGame(seed=mt_seed)
    coach = Coach(g, HashedNet(), dotdict({"numMCTSSims": sims, "cpuct": cpuct, "tempThreshold": temp_threshold}))
    examples = coach.executeEpisode()
    out = [{"key": g.stringRepresentation(b), "v": float(v), "pi": {str(a): float(p).hex() for a, p in enumerate(pi) if p}}
           for b, pi, v in examples]
    return {"mt_seed": mt_seed, "sims": sims, "cpuct": cpuct, "temp_threshold": temp_threshold, "examples": out,
            "rng_after": int(np.random.randint(0, 2 ** 31)), "py_rng_after": random.random().hex()}


def main():
    cases = [
        run_case("uniform_s25", mcts_oracle.uniform_evaluator, 25, 1.5, 0, 0, 15),
        run_case("hashed_s30", mcts_oracle.hashed_evaluator, 30, 1.5, 5, 42, 15),
        run_case("hashed_s64_temp0", lambda b: mcts_oracle.hashed_evaluator(b, 7), 64, 1.1, 9, 1000003, 3, max_plies=14),
        run_case("hashed_s200_late", lambda b: mcts_oracle.hashed_evaluator(b, 3), 200, 2.0, 11, 77, 100, max_plies=6),
    ]
    coach = [coach_episode_case(0, 20, 1.5, 15, 3, 9), coach_episode_case(4, 12, 1.0, 4, 8, 123)]
    with open(os.path.join(HERE, "mcts_golden.json"), "w") as f:
        plain = [coach_plain_case(11, 20, 1.5, 15), coach_plain_case(12, 16, 1.0, 5)]
        json.dump({"numpy": np.__version__, "cases": cases, "coach": coach, "coach_plain": plain}, f, separators=(",", ":"))
    for c in coach:
        print("coach", c["mt_seed"], len(c["examples"]), "examples", c["examples"][-1]["v"], c["rng_after"])
    for c in cases:
        t = c["trace"]
        print(c["name"], "plies", len(t), "result", c["result"], "nodes", c["total_nodes"], "evals", c["leaf_evals"],
              "ns[0..3]", [x["ns"] for x in t[:4]], "last counts", list(t[-1]["counts"].items())[:3])


if __name__ == "__main__":
    main()
