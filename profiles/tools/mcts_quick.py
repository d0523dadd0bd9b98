import sys, time, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200 import mcts_bench
from nypc_yacht_auction_b200.mcts import UniformEvaluator, TorchEvaluator, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
for graph in (False, True):
    r = mcts_bench._run_selfplay(torch, dev, 4096, 25, UniformEvaluator(), 24, 6, 1, 0, graph)
    print('uniform graph=%s' % graph, json.dumps(r))
net = YachtPolicyValueNet().to(dev)
ev = FusedYachtEvaluator(net, 16384)
for graph in (False, True):
    r = mcts_bench._run_selfplay(torch, dev, 16384, 100, ev, 4, 4, 2, 0, graph)
    print('nn graph=%s' % graph, json.dumps(r))
