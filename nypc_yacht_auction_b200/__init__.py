"""nypc_yacht_auction_b200 -- B200-native (sm_100a) batched Yacht-Auction environment and MCTS
self-play engine behind the alpha-zero-general plug-in API of iyioon/NYPC-Yacht-Auction.

Only the hot path lives here: CUDA kernels + C ABI under ``csrc/`` /
``include/yacht_b200.h`` and the Python host mirror of the reference's Game / MCTS interface.
"""
from .layout import YachtBoard, pack_state, string_key  # noqa: F401

__all__ = ["YachtBoard", "pack_state", "string_key"]
