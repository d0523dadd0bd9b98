import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import UniformEvaluator, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
mode = sys.argv[1]
if mode == 'u':
    sp = BatchedSelfPlay(4096, 25, evaluator=UniformEvaluator(), seed=1, device=dev)
else:
    torch.manual_seed(0)
    net = YachtPolicyValueNet().to(dev)
    sp = BatchedSelfPlay(16384, 100, evaluator=FusedYachtEvaluator(net, 16384), seed=1, device=dev)
    sp.mcts.capture_graph()
for t in range(6):
    sp.play_ply(t)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for t in range(6, 10):
        sp.play_ply(t)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
