import sys
sys.path.insert(0, '/root/repo')
import torch
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator, TorchEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
from nypc_yacht_auction_b200.engine import BatchedYacht
torch.manual_seed(0)
net = YachtPolicyValueNet().cuda().eval()
with torch.no_grad():
    for p in net.parameters():
        if p.ndim == 1:
            p.add_(0.1 * torch.randn_like(p))
env = BatchedYacht(300, 1, 1)
for _ in range(7):
    env.play_ply(masks=None, auto_reset=False)
x = env.features()
ev = FusedYachtEvaluator(net, max_batch=512)
logits, v = ev(x)
tev = TorchEvaluator(net, dtype=torch.bfloat16, fused_logits=True)
tl, tv_ = tev(x)
with torch.no_grad():
    rl, rv = net(x)
print("max|logit|", rl.abs().max().item())
print("fused vs fp32", (logits[:, :3226].float() - rl).abs().max().item(), "torch-bf16 vs fp32", (tl[:, :3226].float() - rl).abs().max().item(),
      "fused vs torch-bf16", (logits[:, :3226].float() - tl[:, :3226].float()).abs().max().item())
print("v err fused", (v - rv.reshape(-1)).abs().max().item(), "torch bf16", (tv_ - rv.reshape(-1)).abs().max().item())
sm = lambda t: torch.softmax(t, 1)
print("TV fused", 0.5*(sm(logits[:, :3226].float()) - sm(rl)).abs().sum(1).max().item(), "TV torch", 0.5*(sm(tl[:, :3226].float()) - sm(rl)).abs().sum(1).max().item())
