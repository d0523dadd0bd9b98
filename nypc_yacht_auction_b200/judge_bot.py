"""Judge-protocol client on top of the engine: the NYPC stdin/stdout protocol
READY / ROLL / GET / SCORE / SET / FINISH (/root/reference/INSTRUCTION.md:76-92) that the reference serves with
yacht/submission/agent.py:564-636, here as a thin host of the batched searcher: every decision is one
``BatchedMCTS`` search (n = 1 per session; ``EngineMover`` takes any number of sessions' boards at once) from the
canonical board with "me" as player 1, or the scripted device player (``ya_greedy_action``).

Two layers:

* ``JudgeSession`` -- the protocol state machine, pure host code (no GPU): tracks both players' dice, categories and
  EXACT integer scores from the judge's notifications, builds the board a decision is searched from, turns an engine
  action into ``BID g x`` / ``PUT c d1d2d3d4d5``.
* ``EngineMover`` -- ``choose(boards) -> actions`` on the GPU.  No CPU fallback: without the CUDA library it raises.

Real-valued bids.  The official game allows any integer bid 0..100000 (INSTRUCTION.md:25) while the engine's action
space, like the reference's training environment, has 101 levels 0..50000 step 500 (yacht/YachtGame.py:27-32) and stores
``bid_score`` in units of 500.  Mapping: OUR bids are grid points (legal for the judge as they are); the OPPONENT's bid
arrives through ``GET g g0 x0`` already resolved, so it never has to be an engine action -- only the two running
``bid_score`` totals enter the searched board, rounded to the nearest multiple of 500 (ties away from zero, clamped to
the packed field's range) while the session keeps the exact integers for its own score sheet.
"""
from __future__ import annotations

import itertools
import sys

from .layout import PlayerView, YachtBoard, pack_state

CATEGORIES = ("ONE", "TWO", "THREE", "FOUR", "FIVE", "SIX", "CHOICE", "FOUR_OF_A_KIND", "FULL_HOUSE", "SMALL_STRAIGHT",
              "LARGE_STRAIGHT", "YACHT")                      # yacht/YachtGame.py:15-26 == agent.py DiceRule
N_BID = 202
N_SUBSET = 252
SUBSETS = tuple(itertools.combinations(range(10), 5))         # lexicographic, yacht/YachtGame.py:35
BID_STEP, BID_LEVELS, BID_MAX_JUDGE = 500, 101, 100000


def category_points(cat, dice):
    """Score sheet arithmetic of INSTRUCTION.md:45-66 for the session's own bookkeeping (the search itself scores on the
    device).  Histogram form; agrees with score_category (yacht/YachtGame.py:57-108) on all 7776 x 12 inputs (tested)."""
    hist = [0] * 7
    for d in dice:
        hist[d] += 1
    pips = sum(dice)
    if cat < 6:
        return 1000 * (cat + 1) * hist[cat + 1]
    if cat == 6:
        return 1000 * pips
    if cat == 7:
        return 1000 * pips if max(hist) >= 4 else 0
    if cat == 8:
        return 1000 * pips if (2 in hist[1:] or 5 in hist[1:]) and (3 in hist[1:] or 5 in hist[1:]) else 0
    present = [h > 0 for h in hist]
    if cat == 9:
        return 15000 if any(all(present[s:s + 4]) for s in (1, 2, 3)) else 0
    if cat == 10:
        return 30000 if any(all(present[s:s + 5]) for s in (1, 2)) else 0
    return 50000 if 5 in hist[1:] else 0


def quantise_bid_score(x):
    """Exact running bid total -> the nearest value the packed state can hold (multiple of 500, 13-bit signed)."""
    q = (abs(int(x)) + BID_STEP // 2) // BID_STEP
    q = q if x >= 0 else -q
    return BID_STEP * max(-4096, min(4095, q))


class ProtocolError(ValueError):
    pass


class _Sheet:
    """One player's holdings as the judge has announced them (exact integers)."""

    def __init__(self):
        self.carry = []
        self.used_mask = 0
        self.cat_scores = [0] * 12
        self.bid_score = 0

    def total(self):
        basic = sum(self.cat_scores[:6])
        return sum(self.cat_scores) + (35000 if basic >= 63000 else 0) + self.bid_score

    def view(self):
        return PlayerView(list(self.carry), self.used_mask, list(self.cat_scores), quantise_bid_score(self.bid_score))

    def put(self, cat, dice):
        if (self.used_mask >> cat) & 1:
            raise ProtocolError("category %s already used" % CATEGORIES[cat])
        left = list(self.carry)
        for d in dice:                                         # dice are named by value: take the first matching die each
            if d not in left:
                raise ProtocolError("die %d is not held (holding %r)" % (d, self.carry))
            left.remove(d)
        self.carry = left
        self.used_mask |= 1 << cat
        self.cat_scores[cat] = category_points(cat, dice)


class _Shape:
    """Attribute bag in the reference's YachtState shape for layout.pack_state."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class JudgeSession:
    """One game against the judge.  ``handle(line)`` returns the reply line (without newline) or None; ``mover`` is any
    object with ``choose(boards) -> list[int]`` (EngineMover on the GPU, a stub in the CPU transcript test)."""

    def __init__(self, mover):
        self.mover = mover
        self.me, self.opp = _Sheet(), _Sheet()
        self.round_no = 1
        self.roll_a, self.roll_b = [], []
        self.my_bid = None                                     # (target, amount) of the bid awaiting its GET
        self.scored = [False, False]                           # me, opponent in the current round
        self.finished = False
        self.decisions = 0

    # ------------------------------------------------------------------ boards handed to the search
    def board(self, phase):
        """Canonical board of the pending decision: I am player 1 and to move; no bid is on the table (the opponent's
        bid is secret until resolved); rolls are the current round's bundles (kept through the score phase exactly like
        the reference's state keeps them, yacht/YachtGame.py:290-302)."""
        return pack_state(_Shape(round_no=self.round_no, phase=phase, rollA=list(self.roll_a), rollB=list(self.roll_b),
                                 p1_bid=None, p2_bid=None, p1=self.me.view(), p2=self.opp.view()))

    # ------------------------------------------------------------------ protocol
    def handle(self, line):
        parts = line.split()
        if not parts:
            return None
        cmd, args = parts[0], parts[1:]
        if cmd == "READY":
            return "OK"
        if cmd == "ROLL":
            return self._on_roll(*self._args(cmd, args, 2))
        if cmd == "GET":
            return self._on_get(*self._args(cmd, args, 3))
        if cmd == "SCORE":
            self._args(cmd, args, 0)
            return self._on_score()
        if cmd == "SET":
            return self._on_set(*self._args(cmd, args, 2))
        if cmd == "FINISH":
            self.finished = True
            return None
        raise ProtocolError("Invalid command: %s" % cmd)

    @staticmethod
    def _args(cmd, args, n):
        if len(args) != n:
            raise ProtocolError("%s takes %d argument(s), got %r" % (cmd, n, args))
        return args

    @staticmethod
    def _dice(text):
        if len(text) != 5 or any(c not in "123456" for c in text):
            raise ProtocolError("expected five dice 1-6, got %r" % text)
        return [int(c) for c in text]

    def _on_roll(self, a, b):
        if self.round_no > 12:
            raise ProtocolError("ROLL in round %d (round 13 has no bidding)" % self.round_no)
        self.roll_a, self.roll_b = self._dice(a), self._dice(b)
        action = int(self.mover.choose([self.board(0)])[0])
        self.decisions += 1
        if not 0 <= action < N_BID:                            # a searcher can only return legal moves; belt and braces
            action = 0
        target, amount = "AB"[action // BID_LEVELS], BID_STEP * (action % BID_LEVELS)
        self.my_bid = (target, amount)
        return "BID %s %d" % (target, amount)

    def _on_get(self, got, opp_target, opp_amount):
        if self.my_bid is None:
            raise ProtocolError("GET without a pending BID")
        if got not in ("A", "B") or opp_target not in ("A", "B"):
            raise ProtocolError("bundles are A or B")
        try:
            x0 = int(opp_amount)
        except ValueError:
            raise ProtocolError("opponent bid must be an integer, got %r" % opp_amount)
        if not 0 <= x0 <= BID_MAX_JUDGE:
            raise ProtocolError("opponent bid %d outside 0..%d" % (x0, BID_MAX_JUDGE))
        mine, theirs = (self.roll_a, self.roll_b) if got == "A" else (self.roll_b, self.roll_a)
        self.me.carry.extend(mine)
        self.opp.carry.extend(theirs)
        target, amount = self.my_bid
        self.me.bid_score += -amount if target == got else amount          # INSTRUCTION.md:33-35
        opp_got = "B" if got == "A" else "A"
        self.opp.bid_score += -x0 if opp_target == opp_got else x0
        self.my_bid = None
        if self.round_no == 1:                                 # round 1 has no scoring phase
            self._next_round()
        return None

    def _on_score(self):
        if len(self.me.carry) < 5 or self.scored[0]:
            raise ProtocolError("SCORE without five dice to place")
        action = int(self.mover.choose([self.board(1)])[0])
        self.decisions += 1
        cat, sub = divmod(action - N_BID, N_SUBSET)
        positions = SUBSETS[sub] if 0 <= action - N_BID < 12 * N_SUBSET else ()
        if not positions or max(positions) >= len(self.me.carry) or (self.me.used_mask >> cat) & 1:
            cat = next(c for c in range(12) if not (self.me.used_mask >> c) & 1)   # never forfeit on a bad index
            positions = (0, 1, 2, 3, 4)
        dice = [self.me.carry[i] for i in positions]
        self.me.put(cat, dice)
        self._scored(0)
        return "PUT %s %s" % (CATEGORIES[cat], "".join(str(d) for d in dice))

    def _on_set(self, name, dice):
        if name not in CATEGORIES:
            raise ProtocolError("unknown category %r" % name)
        self.opp.put(CATEGORIES.index(name), self._dice(dice))
        self._scored(1)
        return None

    def _scored(self, who):
        self.scored[who] = True
        if all(self.scored) and self.round_no <= 12:
            self._next_round()

    def _next_round(self):
        self.round_no += 1
        self.scored = [False, False]

    def totals(self):
        return self.me.total(), self.opp.total()


class EngineMover:
    """Decisions on the GPU.  policy "mcts": ``num_sims`` simulations of the batched searcher per decision from a fresh
    tree, most visited move (ties broken uniformly by the engine's Philox stream) -- with a network evaluator
    (``mcts.FusedYachtEvaluator`` of a YachtNNet checkpoint) or the uniform prior; policy "greedy": the reference's
    GreedyYachtPlayer on device (``ya_greedy_action``)."""

    def __init__(self, num_sims=200, evaluator=None, policy="mcts", cpuct=1.5, seed=0, device="cuda", max_boards=1):
        import torch
        from .engine import BatchedYacht
        from .mcts import BatchedMCTS
        if policy not in ("mcts", "greedy"):
            raise ValueError("policy must be 'mcts' or 'greedy'")
        self.torch = torch
        self.policy = policy
        self.n = int(max_boards)
        self.env = BatchedYacht(self.n, seed=seed, device=device)          # raises without CUDA: no CPU fallback
        self.mcts = BatchedMCTS(self.env, int(num_sims), cpuct, evaluator, temp_threshold=0) if policy == "mcts" else None
        self.calls = 0

    def choose(self, boards):
        """Canonical boards (me = player 1, to move) -> one legal action each."""
        torch, env = self.torch, self.env
        boards = list(boards)
        if not 0 < len(boards) <= self.n:
            raise ValueError("between 1 and %d boards per call" % self.n)
        k = len(boards)
        filler = boards + [boards[0]] * (self.n - k)
        env.load_boards(filler, players=[1] * self.n)
        env.ply.fill_(self.calls)                              # a fresh tie-break / sampling stream per decision
        self.calls += 1
        if self.policy == "greedy":
            actions = env.greedy_actions()
        else:
            self.mcts.pool.reset()
            self.mcts.search()
            self.mcts.root_counts()
            actions = self.mcts.pick_actions()
            self.mcts.check_errors()
        return [int(a) for a in actions[:k].cpu().tolist()]


def load_evaluator(checkpoint, device="cuda", max_batch=1, precision="fp16"):
    """A checkpoint written by NNetWrapper.save_checkpoint (yacht/NNet.py:198-205: {"state_dict": ...}) or a bare state
    dict -> FusedYachtEvaluator.  hidden / nblocks are read from the tensors' shapes."""
    import torch
    from .mcts import FusedYachtEvaluator
    from .nnet import YachtPolicyValueNet
    blob = torch.load(checkpoint, map_location="cpu", weights_only=False)
    sd = blob.get("state_dict", blob) if isinstance(blob, dict) else blob
    hidden = sd["inp.0.weight"].shape[0]
    nblocks = len({k.split(".")[1] for k in sd if k.startswith("blocks.")})
    net = YachtPolicyValueNet(input_len=sd["inp.0.weight"].shape[1], action_size=sd["pi_head.2.weight"].shape[0],
                              hidden=hidden, nblocks=nblocks)
    net.load_state_dict(sd)
    return FusedYachtEvaluator(net.to(device).eval(), max_batch, precision=precision)


def serve(mover, stdin=None, stdout=None):
    """The loop of agent.py:564-636: one command per line, replies flushed."""
    stdin, stdout = stdin or sys.stdin, stdout or sys.stdout
    session = JudgeSession(mover)
    for line in stdin:
        reply = session.handle(line.strip())
        if reply is not None:
            stdout.write(reply + "\n")
            stdout.flush()
        if session.finished:
            break
    return session


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="NYPC Yacht-Auction judge-protocol bot on the B200 engine")
    ap.add_argument("--policy", default="mcts", choices=["mcts", "greedy"])
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--checkpoint", default=None, help="YachtNNet checkpoint (yacht/NNet.py save_checkpoint); default: uniform prior")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    evaluator = load_evaluator(args.checkpoint, precision=args.precision) if args.checkpoint else None
    try:
        serve(EngineMover(args.sims, evaluator, args.policy, seed=args.seed))
    except ProtocolError as exc:
        print(str(exc), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
