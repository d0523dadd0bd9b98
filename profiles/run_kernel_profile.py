"""Small driver for ncu captures of the secondary kernels (scoring enumeration, greedy player, MCTS).
    ncu --set full --clock-control none --import-source on -k regex:<name> ... python profiles/run_kernel_profile.py <what>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nypc_yacht_auction_b200.coach import BatchedSelfPlay  # noqa: E402
from nypc_yacht_auction_b200.engine import BatchedYacht  # noqa: E402
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator, UniformEvaluator  # noqa: E402
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet  # noqa: E402

what = sys.argv[1]
dev = torch.device("cuda", 0)
if what == "env":
    env = BatchedYacht(65536, seed=0)
    for _ in range(4):
        env.play_ply(masks=None, auto_reset=False)
    table = torch.empty((65536, 12, 252), dtype=torch.uint8, device=dev)
    for _ in range(3):
        env.enumerate_scores(out=table)
        env.greedy_actions()
elif what == "uniform":
    sp = BatchedSelfPlay(4096, 25, evaluator=UniformEvaluator(), seed=1, device=dev, record_examples=False)
    for t in range(8):
        sp.play_ply(t)
else:
    torch.manual_seed(0)
    net = YachtPolicyValueNet().to(dev)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    sp = BatchedSelfPlay(n, 100, evaluator=FusedYachtEvaluator(net, n), seed=1, device=dev, record_examples=False)
    for t in range(5):
        sp.play_ply(t)
torch.cuda.synchronize()
print("ok", what)
