"""Search-level effect of the reduced-precision leaf evaluator (north_star: "MCTS visit counts and policies match the
reference driven by the same evaluator within a stated float tolerance").

The tree kernels are bit-exact against MCTS.py for a given (pi, v) (tests/test_mcts_gpu.py).  What remains is the
evaluator's arithmetic: the reference's CUDA predict is fp16 autocast (yacht/NNet.py:186-193), its CPU predict float32.
Here the SAME search kernels are driven from the SAME roots by (a) a float32 PyTorch forward of the module and (b) the
hand-written tcgen05 forward in fp16 and in bf16, and the visit distributions are compared.  The measured numbers are
printed, written to gpurun_out/search_tolerance.json, and bounded by the tolerance DESIGN.md states.

A second test pins the float32 route itself: the oracle's restatement of MCTS.py driven by the float32 module on the
CPU (batch 1, exactly like NNetWrapper.predict with cuda=False) against the CUDA kernels fed with the same host
forward -- exact visit counts and moves.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mcts_oracle
from oracle import yacht_rules as yr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Stated tolerance (DESIGN.md "Float semantics of the search"), 100 simulations from 1,024 mid-game roots:
#   fp16 forward vs float32 forward: mean L1 distance of the root visit distributions <= 0.03, the float32 run's most
#   visited move is a most-visited move of the fp16 run for >= 98 % of the roots
#   (measured 0.005-0.013 and 99.0-99.6 %); bf16: <= 0.12 and >= 90 % (measured 0.030-0.072 and 92.8-97.6 %).
TOL = {"fp16": (0.03, 0.98), "bf16": (0.12, 0.90)}


def _roots(n, plies, seed, base):
    from nypc_yacht_auction_b200.engine import BatchedYacht
    env = BatchedYacht(n, seed=seed, game_base=base)
    for _ in range(plies):
        env.play_ply(masks=None, auto_reset=False)                    # deterministic (Philox): every call gives the same roots
    return env


def _search_counts(env, sims, evaluator):
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    m = BatchedMCTS(env, sims, 1.5, evaluator=evaluator)
    m.search()
    m.check_errors()
    counts, visits = m.root_counts()
    return counts.clone(), visits.clone()


def _compare(c_ref, c_other):
    p = c_ref.double() / c_ref.sum(1, keepdim=True).double()
    q = c_other.double() / c_other.sum(1, keepdim=True).double()
    l1 = (p - q).abs().sum(1)
    top_ref = c_ref.argmax(1, keepdim=True)
    agree = (c_other.gather(1, top_ref).squeeze(1) == c_other.max(1).values).double()
    same = (c_ref == c_other).all(1).double()
    return {"mean_l1": float(l1.mean()), "max_l1": float(l1.max()), "top_move_agreement": float(agree.mean()),
            "identical_count_vectors": float(same.mean())}


@pytest.mark.parametrize("sharpen", [1.0, 4.0], ids=["random_init", "peaked_prior"])
def test_reduced_precision_forward_changes_search_within_stated_tolerance(sharpen):
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    from torch_evaluator import TorchEvaluator
    torch.manual_seed(0)
    net = YachtPolicyValueNet().cuda().eval()
    with torch.no_grad():                      # sharpen > 1: a peaked prior and decisive values, as a trained net gives
        net.pi_head[2].weight.mul_(sharpen)
        net.v_head[4].weight.mul_(sharpen)
    n, sims = 1024, 100
    report = {}
    for plies in (9, 22):                      # a bid root (202 moves) and a score root (ten dice, 252 x open categories)
        c32, v32 = _search_counts(_roots(n, plies, 3, 100000), sims, TorchEvaluator(net))
        assert bool((c32.sum(1) == v32).all()) and int(v32.min()) == sims - 1
        again, _ = _search_counts(_roots(n, plies, 3, 100000), sims, TorchEvaluator(net))
        assert torch.equal(c32, again)         # the float32 route is deterministic: differences below are the evaluator's
        for precision in ("fp16", "bf16"):
            c16, v16 = _search_counts(_roots(n, plies, 3, 100000), sims, FusedYachtEvaluator(net, n, precision=precision))
            assert torch.equal(v16, v32)
            r = _compare(c32, c16)
            report["%s_ply%d" % (precision, plies)] = r
            max_l1, min_agree = TOL[precision]
            assert r["mean_l1"] <= max_l1 and r["top_move_agreement"] >= min_agree, (precision, plies, r)
    print("search tolerance (sharpen %g): %s" % (sharpen, json.dumps(report)))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "search_tolerance_%s.json" % ("peaked" if sharpen > 1 else "random_init")), "w") as f:
            json.dump({"sims": sims, "roots": n, "sharpen": sharpen, "vs": "float32 torch forward, same kernels, same roots",
                       "results": report}, f, indent=1)


def test_float32_module_on_cpu_oracle_vs_kernels_exact():
    """8 games x 10 plies x 16 sims: oracle MCTS + the float32 module evaluated on the CPU one leaf at a time (what
    NNetWrapper.predict does without CUDA) against the CUDA tree kernels fed by the same host forward.  Exact."""
    from nypc_yacht_auction_b200.engine import BatchedYacht
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    from test_mcts_gpu import HostEvaluator
    torch.manual_seed(5)
    cpu_net = YachtPolicyValueNet().eval()

    def predict(board):                        # yacht/NNet.py:177-195 with cuda=False
        x = torch.from_numpy(yr.features(board)).unsqueeze(0)
        with torch.no_grad():
            logits, v = cpu_net(x)
        return torch.log_softmax(logits, dim=1).exp()[0].numpy().astype(np.float32), np.float32(v.reshape(-1)[0].item())

    n, sims, plies, seed, base = 8, 16, 10, 17, 300
    traces = [mcts_oracle.self_play_game(predict, sims, 1.5, seed, base + g, max_plies=plies)[0] for g in range(n)]
    env = BatchedYacht(n, seed=seed, game_base=base)
    mcts = BatchedMCTS(env, sims, 1.5, evaluator=HostEvaluator(predict, n), want_leaf_states=True)
    for ply in range(plies):
        mcts.search()
        mcts.check_errors()
        counts, _ = mcts.root_counts()
        c = counts.cpu().numpy()
        acts = mcts.pick_actions().cpu().numpy()
        for g in range(n):
            assert {int(a): int(c[g][a]) for a in np.flatnonzero(c[g])} == traces[g][ply]["counts"], (g, ply)
            assert int(acts[g]) == traces[g][ply]["action"]
        env.next_state(mcts.picked)
