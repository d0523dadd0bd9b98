import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.mcts import TorchEvaluator, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
ev = FusedYachtEvaluator(net, 16384)
x = torch.randn(16384, 59, device=dev)
for _ in range(3):
    pi, v = ev(x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        pi, v = ev(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
