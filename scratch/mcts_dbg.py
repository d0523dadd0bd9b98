import sys, json
sys.path.insert(0, '/root/repo')
import torch
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import TorchEvaluator, UniformEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
n = int(sys.argv[1]); sims = int(sys.argv[2]); mb = float(sys.argv[3])
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
ev = TorchEvaluator(net, autocast_dtype=torch.bfloat16) if len(sys.argv) < 5 else UniformEvaluator()
sp = BatchedSelfPlay(n, sims, evaluator=ev, seed=2, device=dev, arena_mb_per_game=mb, record_examples=False)
peak_top = 0; peak_nodes = 0
for t in range(48):
    sp.play_ply(t)
    top = sp.mcts.pool.meta[:, 1].max().item(); nodes = sp.mcts.pool.meta[:, 0].max().item()
    avg = sp.mcts.pool.meta[:, 1].float().mean().item()
    peak_top = max(peak_top, top); peak_nodes = max(peak_nodes, nodes)
    print(t, "arena words max", top, "avg", int(avg), "nodes max", nodes, "err", hex(sp.mcts.err_flag.item()))
print("peak arena MB", peak_top * 4 / 2**20, "peak nodes", peak_nodes)
