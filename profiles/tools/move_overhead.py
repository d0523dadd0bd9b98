"""Per-move device time outside the search itself (features, root statistics, move sampling, transition)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.coach import BatchedSelfPlay
from nypc_yacht_auction_b200.mcts import UniformEvaluator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sp = BatchedSelfPlay(n, sims, evaluator=UniformEvaluator(), seed=1, device="cuda", record_examples=True)
env, m = sp.env, sp.mcts
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000, r
acc = {}
for t in range(16):
    steps = [("features", lambda: env.features(out=sp.ex_features[t])), ("search", m.search), ("root_counts", m.root_counts),
             ("root_sparse", lambda: m.root_sparse(sp.ex_actions[t], sp.ex_counts[t], sp.ex_overflow)), ("pick", m.pick_actions)]
    for name, fn in steps:
        us, _ = timed(fn)
        if t >= 4: acc[name] = acc.get(name, 0.0) + us / 12
    us, _ = timed(lambda: env.next_state(m.picked, check=False))
    if t >= 4: acc["next_state"] = acc.get("next_state", 0.0) + us / 12
print("n=%d sims=%d per move (us):" % (n, sims), {k: round(v, 1) for k, v in acc.items()}, "total", round(sum(acc.values()), 1))
