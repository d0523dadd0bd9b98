"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's MCTS (/root/reference/MCTS.py) on top of oracle/yacht_rules.py.
It keeps the reference's *arithmetic* expression by expression (numpy scalar types decide
whether a step is float32 or Python double, NEP 50), its state-keyed tables (transpositions
merge, the tree persists across moves), its three leaf fallbacks, the lowest-index tie-break
and the un-negated ``0`` returned from dead-end revisits, but is written
as an explicit per-node record instead of six dicts.

In-search randomness (the reference re-rolls dice inside search, quirk Q1) is taken from the
Philox protocol: draw event (game, episode, ply of the root, TAG_SEARCH, depth, sim).

Pinned by tests/golden/mcts_golden.json, produced by tests/golden/make_golden_mcts.py from the
unmodified reference MCTS.py + YachtGame.py with the same draws injected.
"""
from __future__ import annotations

import math

import numpy as np

from . import philox
from . import yacht_rules as yr

EPS = 1e-8            # MCTS.py:6


class Node:
    __slots__ = ("prior", "legal", "visits", "n_edge", "q_edge")

    def __init__(self, prior, legal):
        self.prior = prior            # float32[3226]          (Ps, MCTS.py:23)
        self.legal = legal            # uint8[3226]            (Vs, MCTS.py:26)
        self.visits = 0               # Ns, MCTS.py:22
        self.n_edge = {}              # Nsa, MCTS.py:21
        self.q_edge = {}              # Qsa, MCTS.py:20 (np.float32 or Python float, as the reference)


def masked_prior(pi, legal):
    """MCTS.py:88-113."""
    p = pi * legal
    total = np.sum(p)
    if total > 0:
        p /= total
        return p
    p = p + legal
    total = np.sum(p)
    if total > 0:
        p /= total
        return p
    p = legal.astype(np.float32)
    if np.sum(p) > 0:                 # unreachable (same test as above); kept for the record
        p /= np.sum(p)
        return p
    p = np.zeros_like(legal, dtype=np.float32)
    p[0] = 1.0
    return p


class TreeSearch:
    """evaluator(board) -> (pi float32[3226], v numpy.float32)   (NeuralNet.predict, NeuralNet.py:27-37)"""

    def __init__(self, evaluator, num_sims, cpuct, seed=0, game_id=0, episode=0):
        self.evaluator = evaluator
        self.num_sims = num_sims
        self.cpuct = cpuct
        self.seed, self.game_id, self.episode = seed, game_id, episode
        self.nodes = {}
        self.terminal = {}            # Es, MCTS.py:25
        self.leaf_evals = 0

    # -- MCTS.py:28-54 ------------------------------------------------------------------
    def root_counts(self, board, ply):
        for sim in range(self.num_sims):
            self._visit(board, ply, sim, 0)
        node = self.nodes[yr.key(board)]
        return np.array([node.n_edge.get(a, 0) for a in range(yr.N_ACTION)], dtype=np.int64)

    def action_probs(self, board, ply, temp=1):
        counts = self.root_counts(board, ply)
        if temp == 0:
            best = np.flatnonzero(counts == counts.max())
            return counts, best
        total = float(sum(float(c) for c in counts))
        return counts, [float(c) / total for c in counts]

    # -- MCTS.py:56-164 -----------------------------------------------------------------
    def _visit(self, board, ply, sim, depth):
        k = yr.key(board)
        if k not in self.terminal:
            self.terminal[k] = yr.outcome(board, 1)
        if self.terminal[k] != 0:
            return -self.terminal[k]

        node = self.nodes.get(k)
        if node is None:
            pi, v = self.evaluator(board)
            self.leaf_evals += 1
            legal = yr.legal_mask(board, 1)
            self.nodes[k] = Node(masked_prior(pi, legal), legal)
            return -v

        best_u = -float("inf")
        best_a = -1
        for a in np.flatnonzero(node.legal):          # ascending, strict '>' keeps the lowest index
            a = int(a)
            if a in node.q_edge:
                u = node.q_edge[a] + self.cpuct * node.prior[a] * math.sqrt(node.visits) / (1 + node.n_edge[a])
            else:
                u = self.cpuct * node.prior[a] * math.sqrt(node.visits + EPS)
            if u > best_u:
                best_u = u
                best_a = a
        if best_a == -1:
            return 0                                   # MCTS.py:138-147: dead end, not negated, no update

        draw = philox.Draw(self.seed, self.game_id, self.episode, ply, philox.TAG_SEARCH, depth, sim)
        nxt, who = yr.next_state(board, 1, best_a, draw)
        nxt = yr.canonical(nxt, who)
        v = self._visit(nxt, ply, sim, depth + 1)

        if best_a in node.q_edge:
            node.q_edge[best_a] = (node.n_edge[best_a] * node.q_edge[best_a] + v) / (node.n_edge[best_a] + 1)
            node.n_edge[best_a] += 1
        else:
            node.q_edge[best_a] = v
            node.n_edge[best_a] = 1
        node.visits += 1
        return -v


# ---------------------------------------------------------------------- deterministic evaluators
def uniform_evaluator(board):
    """BASELINE.json configs[2]: uniform prior, no NN."""
    return np.full(yr.N_ACTION, 1.0 / yr.N_ACTION, dtype=np.float32), np.float32(0.0)


def _mix(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def hashed_evaluator(board, salt=0):
    """A deterministic, platform-independent pseudo-network: integer hashes of the feature row
    give integer weights in [1, 1024]; pi = w / sum(w) (sum exact in float32), v = (h % 2001 - 1000) / 1000.
    Only exactly-rounded float32 operations are used, so every platform produces the same bits."""
    feat = yr.features(board)
    h = salt & 0xFFFFFFFF
    for word in np.frombuffer(feat.tobytes(), dtype=np.uint32):
        h = _mix(h ^ int(word))
    idx = np.arange(yr.N_ACTION, dtype=np.uint64)
    x = (idx * np.uint64(0x9E3779B1) + np.uint64(h)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    w = ((x >> np.uint64(7)) % np.uint64(1024) + np.uint64(1)).astype(np.float32)
    # a peaked component so the search goes deep: a few actions get 64x weight
    w[(x % np.uint64(37)) == 0] *= np.float32(64.0)
    total = np.float32(w.astype(np.float64).sum())           # < 2^24: exact
    pi = (w / total).astype(np.float32)
    v = np.float32((np.float32((_mix(h ^ 0xABCDEF) % 2001)) - np.float32(1000.0)) / np.float32(1000.0))
    return pi, v


# ---------------------------------------------------------------------- self-play driver
def sample_from_counts(counts, word):
    """Inverse-CDF draw over integer visit counts: r = (word * total) >> 32, first action whose
    cumulative count exceeds r (the engine's replacement for np.random.choice(p=pi), Coach.py:65)."""
    total = int(counts.sum())
    r = (word * total) >> 32
    return int(np.searchsorted(np.cumsum(counts), r, side="right"))


def self_play_game(evaluator, num_sims, cpuct, seed, game_id, temp_threshold=15, episode=0, max_plies=None):
    """Coach.executeEpisode (Coach.py:34-72) with Philox draws; returns the per-ply trace."""
    tree = TreeSearch(evaluator, num_sims, cpuct, seed, game_id, episode)
    board = yr.new_game(philox.Draw(seed, game_id, episode, 0, philox.TAG_INIT))
    cur, ply = 1, 0
    trace = []
    while True:
        canon = yr.canonical(board, cur)
        temp = int((ply + 1) < temp_threshold)                       # Coach.py:56-58 (episodeStep is 1-based)
        counts = tree.root_counts(canon, ply)
        word = philox.draw_words(seed, game_id, episode, ply, philox.TAG_ACTION)[3]
        if temp == 0:
            best = np.flatnonzero(counts == counts.max())
            action = int(best[(word * len(best)) >> 32])
        else:
            action = sample_from_counts(counts, word)
        trace.append({"ply": ply, "player": cur, "key": yr.key(canon), "counts": {int(a): int(counts[a]) for a in np.flatnonzero(counts)},
                      "action": action, "nodes": len(tree.nodes)})
        board, cur = yr.next_state(board, cur, action, philox.Draw(seed, game_id, episode, ply, philox.TAG_REAL))
        ply += 1
        r = yr.outcome(board, cur)
        if r != 0 or (max_plies is not None and ply >= max_plies):
            return trace, board, cur, r, tree
