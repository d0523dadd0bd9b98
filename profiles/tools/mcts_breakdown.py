import sys, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200 import _lib
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import TorchEvaluator, UniformEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
n = int(sys.argv[1]); sims = int(sys.argv[2]); uniform = len(sys.argv) > 3
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
ev = UniformEvaluator() if uniform else TorchEvaluator(net, dtype=torch.bfloat16, fused_logits=True)
sp = BatchedSelfPlay(n, sims, evaluator=ev, seed=2, device=dev, record_examples=False)
m = sp.mcts; env = sp.env; lib = m.lib
def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(12):
    acc = {"select": 0.0, "eval": 0.0, "expand": 0.0}
    for sim in range(sims):
        s = _lib.current_stream()
        a0, a1 = ev_pair(); b0, b1 = ev_pair(); c0, c1 = ev_pair()
        a0.record()
        _lib.check(lib.ya_mcts_select(m.pool.ref, _lib.ptr(env.states), env.n, _lib.ptr(env.players), _lib.ptr(env.ply), _lib.ptr(env.episode),
            env.seed, env.game_base, sim, None, m.cpuct, None, _lib.ptr(m.features), _lib.ptr(m.need_eval), None, _lib.ptr(m.err_flag), s), "sel")
        a1.record(); b0.record()
        if not uniform:
            pi, v = ev(m.features)
        b1.record(); c0.record()
        if uniform:
            _lib.check(lib.ya_mcts_expand(m.pool.ref, None, None, 1, ev.p, ev.v, None, _lib.ptr(m.err_flag), s), "exp")
        else:
            _lib.check(lib.ya_mcts_expand_logits(m.pool.ref, _lib.ptr(pi), pi.shape[1], _lib.ptr(v), None, _lib.ptr(m.err_flag), s), "exp")
        c1.record()
        torch.cuda.synchronize()
        acc["select"] += a0.elapsed_time(a1); acc["eval"] += b0.elapsed_time(b1); acc["expand"] += c0.elapsed_time(c1)
    m.root_counts(); a = m.pick_actions(); env.next_state(a, check=False)
    print("ply", t, {k: round(v / sims * 1000, 1) for k, v in acc.items()}, "us per wave; leaves/wave", int(m.need_eval.sum().item()))
m.check_errors()
