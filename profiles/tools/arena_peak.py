import sys; sys.path.insert(0, "/root/repo")
import torch
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
torch.manual_seed(0)
net = YachtPolicyValueNet().cuda().eval()
sp = BatchedSelfPlay(4096, 100, evaluator=FusedYachtEvaluator(net, 4096), seed=3, record_examples=False)
sp.mcts.capture_graph()
peak_top = peak_nodes = 0
for t in range(48):
    sp.play_ply(t)
    meta = sp.mcts.pool.meta
    peak_top = max(peak_top, int(meta[:, 1].max().item())); peak_nodes = max(peak_nodes, int(meta[:, 0].max().item()))
sp.mcts.check_errors()
print("arena words capacity %d, peak top %d (%.1f%%); nodes capacity %d, peak %d" % (sp.mcts.pool.arena_words, peak_top, 100.0 * peak_top / sp.mcts.pool.arena_words, sp.mcts.pool.max_nodes, peak_nodes))
