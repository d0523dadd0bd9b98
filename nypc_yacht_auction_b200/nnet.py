"""Leaf evaluator network: state-dict compatible with the reference's YachtNNet
(/root/reference/yacht/pytorch/YachtNNet.py:8-70; keys inp.0/1, blocks.i.fc1/ln1/fc2/ln2, pi_head.0/2,
v_head.0/2/4) so checkpoints written by yacht/NNet.py:198-205 load unchanged.  This module only HOLDS the weights
(and is the float32 yardstick of the tests): on the self-play path the forward of a whole simulation wave is the one
hand-written tcgen05 kernel csrc/ya_forward.cu, fed from this state dict by mcts.FusedYachtEvaluator.  Dropout is a
no-op in eval mode and is omitted.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, dim)
        self.ln1 = nn.LayerNorm(dim)
        self.fc2 = nn.Linear(dim, dim)
        self.ln2 = nn.LayerNorm(dim)

    def forward(self, x):
        h = self.ln1(F.silu(self.fc1(x)))
        return x + self.ln2(F.silu(self.fc2(h)))


class YachtPolicyValueNet(nn.Module):
    def __init__(self, input_len=59, action_size=3226, hidden=256, nblocks=6):
        super().__init__()
        self.inp = nn.Sequential(nn.Linear(input_len, hidden), nn.LayerNorm(hidden), nn.SiLU(), nn.Identity())
        self.blocks = nn.ModuleList(_Block(hidden) for _ in range(nblocks))
        self.pi_head = nn.Sequential(nn.LayerNorm(hidden), nn.SiLU(), nn.Linear(hidden, action_size))
        self.v_head = nn.Sequential(nn.LayerNorm(hidden), nn.SiLU(), nn.Linear(hidden, 128), nn.SiLU(), nn.Linear(128, 1))
        for m in self.modules():                      # YachtNNet._init, YachtNNet.py:56-60
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
                nn.init.zeros_(m.bias)

    def forward(self, x):
        if x.ndim == 3:
            x = x.squeeze(1)
        h = self.inp(x)
        for blk in self.blocks:
            h = blk(h)
        return self.pi_head(h), torch.tanh(self.v_head(h))

    @staticmethod
    def num_macs(input_len=59, action_size=3226, hidden=256, nblocks=6):
        return input_len * hidden + nblocks * 2 * hidden * hidden + hidden * action_size + hidden * 128 + 128
