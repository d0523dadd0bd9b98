"""Batched, device-resident Yacht-Auction environment: thousands of lock-step games, one
thread (transition) or one CTA slice (legal-mask stream) per game, every operation a
hand-written sm_100a kernel reached through the C ABI (include/yacht_b200.h).

The method names mirror the reference's Game API (Game.py / yacht/YachtGame.py) with a
leading batch dimension; PyTorch only owns the device buffers and the stream.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .layout import boards_to_planes, planes_to_boards

ACTION_SIZE = 3226
FEATURE_SIZE = 59
SCORE_TABLE = 3024

TAG_INIT, TAG_REAL, TAG_ACTION, TAG_SEARCH = 0, 1, 2, 3

STATUS_EXC = {
    1: (ValueError, "Invalid action in BID phase"),
    2: (ValueError, "Invalid action in SCORE phase"),
    3: (RuntimeError, "Invalid phase/state"),
    4: (AssertionError, "bid resolution without both bids"),
    5: (OverflowError, "carry would exceed 10 dice (state outside legal play)"),
}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise _lib.YachtB200Error("no CUDA device: the Yacht-Auction engine has no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.YachtB200Error("device must be a CUDA device, got %r" % (device,))
    return dev


class BatchedYacht:
    """n concurrent games in HBM (32 B each, two uint4 planes) plus per-game side arrays.

    seed / game_base define the Philox streams: game g of this batch is global game
    ``game_base + g`` so results do not depend on how games are sharded over GPUs.
    """

    def __init__(self, n, seed=0, game_base=0, device="cuda"):
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.n = int(n)
        self.seed = int(seed)
        self.game_base = int(game_base)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ya_set_device(self.device.index or 0), "ya_set_device")
        d = self.device
        self.states = torch.zeros((2, self.n, 4), dtype=torch.int32, device=d)
        self.players = torch.ones(self.n, dtype=torch.int8, device=d)
        self.ply = torch.zeros(self.n, dtype=torch.int32, device=d)
        self.episode = torch.zeros(self.n, dtype=torch.int32, device=d)     # reinterpreted as uint32
        self.actions = torch.zeros(self.n, dtype=torch.int32, device=d)
        self.outcome = torch.zeros(self.n, dtype=torch.float32, device=d)
        self.status = torch.zeros(self.n, dtype=torch.int32, device=d)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=d)
        self.reset()

    # ------------------------------------------------------------------ helpers
    def _s(self):
        return _lib.current_stream()

    def reset(self):
        """getInitBoard for every game (YachtGame.py:232-237)."""
        self.ply.zero_()
        _lib.check(self.lib.ya_init_states(_lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(self.ply),
                                           _lib.ptr(self.episode), self.n, self.seed, self.game_base, self._s()),
                   "ya_init_states")

    def load_boards(self, boards, players=None):
        planes = boards_to_planes(boards)
        assert planes.shape[1] == self.n
        self.states.copy_(torch.from_numpy(planes.view(np.int32)))
        if players is not None:
            self.players.copy_(torch.as_tensor(np.asarray(players, dtype=np.int8)))

    def boards(self):
        return planes_to_boards(self.states.cpu().numpy().view(np.uint32))

    def raise_on_status(self, status=None):
        st = (self.status if status is None else status).cpu().numpy()
        bad = np.nonzero(st)[0]
        if len(bad):
            code = int(st[bad[0]])
            exc, msg = STATUS_EXC.get(code, (RuntimeError, "status %#x" % code))
            raise exc("%s (game %d)" % (msg, int(bad[0])))

    # ------------------------------------------------------------------ Game API, batched
    def valid_moves(self, out=None, states=None, players=None):
        """getValidMoves (YachtGame.py:374-406) -> uint8[n, 3226]."""
        st = self.states if states is None else states
        pl = self.players if players is None else players
        n = st.shape[1]
        if out is None:
            out = torch.empty((n, ACTION_SIZE), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.ya_valid_moves(_lib.ptr(st), n, _lib.ptr(pl), _lib.ptr(out), n, self._s()), "ya_valid_moves")
        return out

    def next_state(self, actions, check=True, tag=TAG_REAL, injected=None):
        """getNextState (YachtGame.py:260-372) in place for every game; bumps ply."""
        actions = actions.to(device=self.device, dtype=torch.int32)
        mode = 0 if injected is None else 1
        _lib.check(self.lib.ya_next_state(
            _lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(actions),
            _lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(self.status), self.n,
            mode, _lib.ptr(injected), self.seed, self.game_base, _lib.ptr(self.episode), _lib.ptr(self.ply),
            tag, None, self._s()), "ya_next_state")
        if check:
            self.raise_on_status()
        self.ply += 1
        return self.players

    def game_ended(self, players=None, out=None):
        """getGameEnded (YachtGame.py:408-428) -> float32[n]."""
        pl = self.players if players is None else players
        if out is None:
            out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ya_game_ended(_lib.ptr(self.states), self.n, _lib.ptr(pl), _lib.ptr(out), self.n, self._s()),
                   "ya_game_ended")
        return out

    def canonical(self, out=None):
        """getCanonicalForm (YachtGame.py:430-442) -> int32[2, n, 4]."""
        if out is None:
            out = torch.empty_like(self.states)
        _lib.check(self.lib.ya_canonical_form(_lib.ptr(self.states), self.n, _lib.ptr(self.players),
                                              _lib.ptr(out), self.n, self.n, self._s()), "ya_canonical_form")
        return out

    def features(self, canonical_states=None, out=None):
        """state_to_vec (yacht/NNet.py:50-86) of canonical states -> float32[n, 59]."""
        st = self.canonical() if canonical_states is None else canonical_states
        n = st.shape[1]
        if out is None:
            out = torch.empty((n, FEATURE_SIZE), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ya_features(_lib.ptr(st), n, _lib.ptr(out), n, self._s()), "ya_features")
        return out

    def random_actions(self, out=None):
        """RandomYachtPlayer.play (yacht/YachtPlayers.py:174-183) with the Philox pick."""
        if out is None:
            out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.ya_random_action(_lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(out), self.n,
                                             self.seed, self.game_base, _lib.ptr(self.episode), _lib.ptr(self.ply),
                                             self._s()), "ya_random_action")
        return out

    def enumerate_scores(self, out=None):
        """All 12 x 252 (category, subset) scores / 1000 for the player to move -> uint8[n, 12, 252]."""
        if out is None:
            out = torch.empty((self.n, 12, 252), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.ya_enumerate_scores(_lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(out),
                                                self.n, self._s()), "ya_enumerate_scores")
        return out

    def greedy_actions(self, out=None, raw=None, fallback=True):
        """GreedyYachtPlayer.play (yacht/YachtPlayers.py:186-214) for the player to move of every game."""
        if out is None:
            out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.ya_greedy_action(_lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(out), _lib.ptr(raw),
                                             self.n, 1 if fallback else 0, self.seed, self.game_base, _lib.ptr(self.episode),
                                             _lib.ptr(self.ply), self._s()), "ya_greedy_action")
        return out

    def play_ply(self, masks=None, auto_reset=True):
        """One fused ply under the random-legal policy (Arena.py:49-71 with RandomYachtPlayer):
        mask (optional) + sampled action + transition + outcome, one kernel launch."""
        _lib.check(self.lib.ya_play_ply(
            _lib.ptr(self.states), self.n, _lib.ptr(self.players), _lib.ptr(self.ply), _lib.ptr(self.episode),
            _lib.ptr(self.actions), _lib.ptr(self.outcome), _lib.ptr(masks), _lib.ptr(self.err_flag),
            self.n, self.seed, self.game_base, 1 if auto_reset else 0, self._s()), "ya_play_ply")
        return self.actions, self.outcome

    # ------------------------------------------------------------------ CUDA-graph replay of a full game
    def capture_game_graph(self, masks=None, plies=48, auto_reset=True):
        """Capture `plies` fused plies (one full game for every slot) as ONE CUDA graph: the launch-bound
        inner loop is replayed with a single host call (no per-launch gaps)."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.play_ply(masks=masks, auto_reset=auto_reset)        # warm-up launch outside capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(plies):
                self.play_ply(masks=masks, auto_reset=auto_reset)
        return graph
