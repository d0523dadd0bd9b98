"""Bulk parity: 16,384 games x 100 plies (two full episodes with on-device re-deals) of the fused CUDA ply
against the plain-C oracle (oracle/yacht_oracle.c): every packed state word, every sampled action, the
legal-move count of every mask and every outcome are bit-identical (1.6 M game steps)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def test_fused_ply_vs_c_oracle_bulk():
    from nypc_yacht_auction_b200.engine import BatchedYacht
    n, plies, seed, base = 16384, 100, 2025, 123456
    ref = c_oracle.play_random(n, plies, seed, base, auto_reset=True)
    env = BatchedYacht(n, seed=seed, game_base=base)
    masks = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    codes = {0.0: 0, 1.0: 1, -1.0: -1}
    for t in range(plies):
        before_reset = None
        acts, outcome = env.play_ply(masks=masks, auto_reset=False)
        legal = masks.sum(dim=1, dtype=torch.int32).cpu().numpy()
        st = env.states.cpu().numpy().view(np.uint32)                      # [2, n, 4]
        words = np.concatenate([st[0], st[1]], axis=1)                     # [n, 8]
        assert (words == ref["packed"][t]).all(), t
        assert (acts.cpu().numpy() == ref["actions"][t]).all(), t
        assert (legal == ref["legal"][t]).all(), t
        out = outcome.cpu().numpy()
        code = np.where(np.abs(out) == 1.0, out, np.where(out != 0, 2, 0)).astype(np.int8)
        assert (code == ref["result"][t]).all(), t
        if (out != 0).any():                                               # re-deal exactly like auto_reset does
            assert (out != 0).all()
            env.episode += 1
            env.reset()
    assert int(env.err_flag.item()) == 0
    assert ref["steps"] == n * plies


def test_auto_reset_path_vs_c_oracle():
    """Same comparison with the kernel's own re-deal (auto_reset=1): states after the ply that ends a game
    are the freshly dealt boards of the next episode."""
    from nypc_yacht_auction_b200.engine import BatchedYacht
    n, seed, base = 4096, 7, 99
    ref = c_oracle.play_random(n, 60, seed, base, auto_reset=True)
    env = BatchedYacht(n, seed=seed, game_base=base)
    for t in range(60):
        acts, outcome = env.play_ply(masks=None, auto_reset=True)
        assert (acts.cpu().numpy() == ref["actions"][t]).all(), t
        if t != 47:
            st = env.states.cpu().numpy().view(np.uint32)
            assert (np.concatenate([st[0], st[1]], axis=1) == ref["packed"][t]).all(), t
    assert (env.episode.cpu().numpy() == 1).all()
