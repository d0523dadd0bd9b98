import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.engine import BatchedYacht
from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
for groups in (1, 2, 4):
    env = BatchedYacht(16384, seed=2, device=dev)
    m = BatchedMCTS(env, 100, 1.5, evaluator=FusedYachtEvaluator(net, 16384), groups=groups)
    m.capture_graph()
    for t in range(4):
        m.play_ply()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(4):
        m.play_ply()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    m.check_errors()
    print("groups", groups, "ms", round(ms, 1), "sims/s %.3e" % (16384 * 100 * 4 / (ms * 1e-3)))
    del m, env
    torch.cuda.empty_cache()
