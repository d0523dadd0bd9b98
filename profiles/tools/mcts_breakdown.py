"""Per-wave device time of select / leaf evaluator / expand during batched AlphaZero self-play (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200 import _lib
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator, UniformEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device('cuda', 0)
n = int(sys.argv[1]); sims = int(sys.argv[2]); uniform = len(sys.argv) > 3 and sys.argv[3] == "uniform"
arena_mb = float(sys.argv[4]) if len(sys.argv) > 4 else None
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev)
ev = UniformEvaluator() if uniform else FusedYachtEvaluator(net, n)
sp = BatchedSelfPlay(n, sims, evaluator=ev, seed=2, device=dev, record_examples=False, arena_mb_per_game=arena_mb)
m = sp.mcts; env = sp.env; lib = m.lib; grp = m.groups[0]
def pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(12):
    acc = {"select": 0.0, "eval": 0.0, "expand": 0.0}
    m.begin_move() if hasattr(m, "begin_move") else None
    for sim in range(sims):
        s = _lib.current_stream()
        a0, a1 = pair(); b0, b1 = pair(); c0, c1 = pair()
        a0.record()
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
            env.seed, env.game_base, sim, None, None, m.cpuct, None, _lib.ptr(grp.features), _lib.ptr(grp.need_eval),
            _lib.ptr(m.leaf_states), m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc), _lib.ptr(m.err_flag), s), "sel")
        a1.record(); b0.record()
        if not uniform:
            pi, v = ev(grp.features, grp.need_eval, m.leaf_states)
        b1.record(); c0.record()
        if uniform:
            _lib.check(lib.ya_mcts_expand(grp.ref, None, None, 1, ev.p, ev.v, None, _lib.ptr(m.err_flag), s), "exp")
        else:
            _lib.check(lib.ya_mcts_expand_logits(grp.ref, _lib.ptr(pi), pi.shape[1], _lib.ptr(getattr(ev, "last_row_max", None)), _lib.ptr(v), None, _lib.ptr(m.err_flag), s), "exp")
        c1.record()
        torch.cuda.synchronize()
        acc["select"] += a0.elapsed_time(a1); acc["eval"] += b0.elapsed_time(b1); acc["expand"] += c0.elapsed_time(c1)
    m.root_counts(); a = m.pick_actions(); env.next_state(a, check=False)
    print("ply", t, {k: round(v / sims * 1000, 1) for k, v in acc.items()}, "us per wave; leaves/wave", int((grp.need_eval == 1).sum().item()))
m.check_errors()
