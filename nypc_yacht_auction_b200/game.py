"""Drop-in ``YachtGame``: the alpha-zero-general Game API of the reference
(/root/reference/Game.py:14-113, concrete yacht/YachtGame.py:210-474) with every rule evaluated by
the CUDA kernels through the C ABI.  Boards are ``YachtBoard`` value objects (32 packed bytes +
the reference's attribute names), so Coach / Arena / MCTS / state_to_vec / YachtPlayers code written
against the reference runs unchanged on top of this class.

Randomness: like the reference, dice and tie-breaks of *real* moves come from the two module-level
hooks ``roll_five`` / ``tiebreak_uniform`` (reference: yacht/YachtGame.py:154-159), which draw from
the global numpy / ``random`` generators seeded by ``YachtGame(seed)``; they are called exactly
when the reference would call them (the kernel reports which draws a transition needs) and their
values are injected into the kernel, so a seeded game is bit-identical to the reference's.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib
from .engine import ACTION_SIZE, FEATURE_SIZE, STATUS_EXC
from .layout import YachtBoard, boards_to_planes, pack_state, planes_to_boards, string_key

NEED_TIE, NEED_ROLLS = 0x100, 0x200


def roll_five():
    """yacht/YachtGame.py:154-155."""
    return list(np.random.randint(1, 7, size=5))


def tiebreak_uniform():
    """yacht/YachtGame.py:158-159."""
    return random.randint(0, 1)


class YachtGame:
    def __init__(self, seed=None, device="cuda"):
        if seed is not None:                       # yacht/YachtGame.py:222-225
            np.random.seed(seed)
            random.seed(seed)
        if not torch.cuda.is_available():
            raise _lib.YachtB200Error("no CUDA device: YachtGame has no CPU fallback")
        self.device = torch.device(device)
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ya_set_device(self.device.index or 0), "ya_set_device")
        d = self.device
        # batch-of-one device buffers + pinned staging (one H2D / D2H pair per call)
        self.d_state = torch.zeros((2, 1, 4), dtype=torch.int32, device=d)
        self.d_out = torch.zeros((2, 1, 4), dtype=torch.int32, device=d)
        self.d_player = torch.ones(1, dtype=torch.int8, device=d)
        self.d_next = torch.ones(1, dtype=torch.int8, device=d)
        self.d_action = torch.zeros(1, dtype=torch.int32, device=d)
        self.d_status = torch.zeros(1, dtype=torch.int32, device=d)
        self.d_inj = torch.zeros(12, dtype=torch.uint8, device=d)
        self.d_mask = torch.zeros((1, ACTION_SIZE), dtype=torch.uint8, device=d)
        self.d_float = torch.zeros(1, dtype=torch.float32, device=d)
        self.d_feat = torch.zeros((1, FEATURE_SIZE), dtype=torch.float32, device=d)
        self.d_scores = torch.zeros((1, 12, 252), dtype=torch.uint8, device=d)

    # ------------------------------------------------------------------ plumbing
    def _put(self, board, player=1):
        b = pack_state(board)
        self.d_state.copy_(torch.from_numpy(boards_to_planes([b]).view(np.int32)))
        self.d_player.fill_(1 if player == 1 else -1)
        return b

    @staticmethod
    def _stream():
        return _lib.current_stream()

    # ------------------------------------------------------------------ Game API
    def getInitBoard(self):
        """yacht/YachtGame.py:232-237: rollA then rollB through the hook."""
        a = [int(x) for x in roll_five()]
        b = [int(x) for x in roll_five()]
        w1 = sum(d << (3 * i) for i, d in enumerate(a)) | (sum(d << (3 * i) for i, d in enumerate(b)) << 15)
        return YachtBoard((1, w1, 0, 0, 0, 0, 0, 0))

    def getBoardSize(self):
        return (1, 59)                              # yacht/YachtGame.py:239-255

    def getActionSize(self):
        return ACTION_SIZE                          # yacht/YachtGame.py:257-258

    def getNextState(self, board, player, action):
        """yacht/YachtGame.py:260-372 on the GPU; draws injected from the hooks (tie, rollA, rollB order)."""
        self._put(board, player)
        self.d_action.fill_(int(action))
        inj = np.zeros(12, dtype=np.uint8)
        self.d_inj.zero_()
        for _ in range(2):
            _lib.check(self.lib.ya_next_state(
                _lib.ptr(self.d_state), 1, _lib.ptr(self.d_player), _lib.ptr(self.d_action),
                _lib.ptr(self.d_out), 1, _lib.ptr(self.d_next), _lib.ptr(self.d_status), 1,
                1, _lib.ptr(self.d_inj), 0, 0, None, None, 1, None, self._stream()), "ya_next_state")
            status = int(self.d_status.item())
            if not status & (NEED_TIE | NEED_ROLLS):
                break
            if status & NEED_TIE:                   # :522
                inj[0] = int(tiebreak_uniform())
                inj[11] |= 1
            if status & NEED_ROLLS:                 # :298-299 / :358-359
                inj[1:6] = [int(x) for x in roll_five()]
                inj[6:11] = [int(x) for x in roll_five()]
                inj[11] |= 2
            self.d_inj.copy_(torch.from_numpy(inj))
        if status:
            exc, msg = STATUS_EXC.get(status, (RuntimeError, "status %#x" % status))
            raise exc(msg)
        nxt = planes_to_boards(self.d_out.cpu().numpy().view(np.uint32))[0]
        return nxt, int(self.d_next.item())

    def getValidMoves(self, board, player):
        """yacht/YachtGame.py:374-406 -> np.uint8[3226]."""
        self._put(board, player)
        _lib.check(self.lib.ya_valid_moves(_lib.ptr(self.d_state), 1, _lib.ptr(self.d_player), _lib.ptr(self.d_mask), 1,
                                           self._stream()), "ya_valid_moves")
        return self.d_mask[0].cpu().numpy()

    def getGameEnded(self, board, player):
        """yacht/YachtGame.py:408-428 -> 0.0 / +-1.0 / 1e-4 (Python float)."""
        self._put(board, player)
        _lib.check(self.lib.ya_game_ended(_lib.ptr(self.d_state), 1, _lib.ptr(self.d_player), _lib.ptr(self.d_float), 1,
                                          self._stream()), "ya_game_ended")
        r = float(self.d_float.item())
        if r == 0.0 or abs(r) == 1.0:
            return r
        return 1e-4

    def getCanonicalForm(self, board, player):
        """yacht/YachtGame.py:430-442 (player 1 returns the same object, like the reference)."""
        if player == 1:
            return board
        self._put(board, player)
        _lib.check(self.lib.ya_canonical_form(_lib.ptr(self.d_state), 1, _lib.ptr(self.d_player), _lib.ptr(self.d_out), 1, 1,
                                              self._stream()), "ya_canonical_form")
        return planes_to_boards(self.d_out.cpu().numpy().view(np.uint32))[0]

    def getSymmetries(self, board, pi):
        return [(board, pi)]                        # yacht/YachtGame.py:444-446

    def stringRepresentation(self, board):
        return string_key(board)                    # yacht/YachtGame.py:448-467 (same text key)

    def display(self, board):
        pass                                        # yacht/YachtGame.py:472-474

    # ------------------------------------------------------------------ extras on the same kernels
    def stateToVec(self, canonical_board):
        """state_to_vec (yacht/NNet.py:65-86) -> np.float32[59]."""
        self._put(canonical_board, 1)
        _lib.check(self.lib.ya_features(_lib.ptr(self.d_state), 1, _lib.ptr(self.d_feat), 1, self._stream()), "ya_features")
        return self.d_feat[0].cpu().numpy()

    def scoreTable(self, board, player):
        """All 12 x 252 (category, subset) scores in points for `player` (scoring-move enumeration)."""
        self._put(board, player)
        _lib.check(self.lib.ya_enumerate_scores(_lib.ptr(self.d_state), 1, _lib.ptr(self.d_player), _lib.ptr(self.d_scores), 1,
                                                self._stream()), "ya_enumerate_scores")
        return self.d_scores[0].cpu().numpy().astype(np.int32) * 1000
