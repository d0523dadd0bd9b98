"""Game groups on separate streams: do the latency-bound tree kernels of one group overlap the tensor-core forward of another?
    python profiles/tools/groups_sweep.py [games]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.engine import BatchedYacht
from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
for n in [int(a) for a in sys.argv[1:]] or [16384, 37888]:
    for groups, prio in ((1, False), (2, False), (2, True), (3, True), (4, True), (6, True)):
        env = BatchedYacht(n, seed=2, device=dev)
        m = BatchedMCTS(env, 100, 1.5, evaluator=FusedYachtEvaluator(net, n), groups=groups, forward_priority=prio)
        m.capture_graph()
        for t in range(6):
            m.play_ply()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(4):
            m.play_ply()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        m.check_errors()
        print("games", n, "groups", groups, "forward_priority", prio, "ms", round(ms, 1), "sims/s %.3e" % (n * 100 * 4 / (ms * 1e-3)), flush=True)
        del m, env
        torch.cuda.empty_cache()
