"""Times the leaf evaluator alone (16,384 leaves): whole-forward tcgen05 kernel vs trunk kernel + cuBLAS heads vs layer-by-layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
net = YachtPolicyValueNet().to(dev).eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
x = torch.rand((n, 59), device=dev)
for name, kw in (("whole_forward", {}), ("trunk_kernel", dict(whole_forward=False)), ("layerwise", dict(trunk_kernel=False, whole_forward=False))):
    ev = FusedYachtEvaluator(net, n, **kw)
    for _ in range(5):
        ev(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ev(x)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 50
    flops = 2.0 * n * (64 * 256 + 24 * 256 * 256 + 256 * 128 + 256 * 3328)
    print("%-14s n=%d  %.1f us/forward  %.1f TFLOP/s (bf16, padded shapes)" % (name, n, us, flops / us * 1e-6))
