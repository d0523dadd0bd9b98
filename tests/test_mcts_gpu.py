"""Parity of the CUDA MCTS (select / expand / backup kernels through the C ABI) against
(a) golden traces of the reference's own MCTS.py and (b) the oracle, with identical evaluators and
identical injected dice.  Visit counts, chosen actions and node creation are exact; root Q values are
compared bit-for-bit including their numeric type (numpy float32 vs Python float)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mcts_oracle
from oracle import yacht_rules as yr
from conftest import to_oracle_board

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_golden.json")
EVALUATORS = {
    "uniform_s25": mcts_oracle.uniform_evaluator,
    "hashed_s30": mcts_oracle.hashed_evaluator,
    "hashed_s64_temp0": lambda b: mcts_oracle.hashed_evaluator(b, 7),
    "hashed_s200_late": lambda b: mcts_oracle.hashed_evaluator(b, 3),
}


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


class HostEvaluator:
    """Test-only evaluator: runs the oracle's deterministic pseudo-network on the host for every leaf."""
    uniform = False

    def __init__(self, fn, n):
        self.fn = fn
        self.evals = 0
        self.pi = torch.zeros((n, 3226), dtype=torch.float32)
        self.v = torch.zeros(n, dtype=torch.float32)

    def __call__(self, features, need_eval, leaf_states):
        from nypc_yacht_auction_b200.layout import planes_to_boards
        need = need_eval.cpu().numpy()
        boards = planes_to_boards(leaf_states.cpu().numpy().view(np.uint32))
        feats = features.cpu().numpy()
        for g in np.flatnonzero(need):
            ob = to_oracle_board(boards[g])
            assert feats[g].tobytes() == yr.features(ob).tobytes()        # the device feature row of the leaf
            pi, v = self.fn(ob)
            self.pi[g] = torch.from_numpy(pi)
            self.v[g] = float(v)
            self.evals += 1
        return self.pi.cuda(), self.v.cuda()


def _engine(n, seed, base):
    from nypc_yacht_auction_b200.engine import BatchedYacht
    return BatchedYacht(n, seed=seed, game_base=base)


@pytest.mark.parametrize("case", load_cases(), ids=lambda c: c["name"])
def test_batched_mcts_reproduces_reference_traces(case):
    from nypc_yacht_auction_b200.mcts import BatchedMCTS, UniformEvaluator
    from nypc_yacht_auction_b200.layout import string_key
    env = _engine(1, case["seed"], case["game"])
    if case["name"].startswith("uniform"):
        ev = UniformEvaluator()
    else:
        ev = HostEvaluator(EVALUATORS[case["name"]], 1)
    mcts = BatchedMCTS(env, case["sims"], case["cpuct"], evaluator=ev, temp_threshold=case["temp_threshold"],
                       want_leaf_states=True)
    prev_nodes = 0
    for ref in case["trace"]:
        canon = env.canonical()
        from nypc_yacht_auction_b200.layout import planes_to_boards
        assert string_key(planes_to_boards(canon.cpu().numpy().view(np.uint32))[0]) == ref["key"]
        before = getattr(ev, "evals", 0)
        mcts.search()
        mcts.check_errors()
        counts, visits = mcts.root_counts()
        c = counts[0].cpu().numpy()
        got = {str(int(a)): int(c[a]) for a in np.flatnonzero(c)}
        assert got == ref["counts"], ref["ply"]
        assert int(visits.item()) == ref["ns"]
        if not case["name"].startswith("uniform"):
            assert ev.evals - before == ref["nodes"] - prev_nodes       # same leaves created this move
        prev_nodes = ref["nodes"]
        a = mcts.pick_actions()
        assert int(a.item()) == ref["action"]
        last = ref
        if ref is case["trace"][-1]:
            _, _, q, kind = mcts.root_counts(with_q=True)
            q, kind = q[0].cpu().numpy(), kind[0].cpu().numpy()
            for act, (k, hexval) in ref["q"].items():
                assert {1: "f32", 2: "f64"}[int(kind[int(act)])] == k
                assert float(q[int(act)]).hex() == hexval
        env.next_state(a)
    assert string_key(env.boards()[0]) == case["final_key"]
    assert float(env.game_ended().item()) == case["result"]
    assert int(env.players.item()) == case["final_player"]


def test_batched_mcts_many_games_vs_oracle():
    """Several games in one batch, hashed evaluator, compared per ply with the oracle's self-play."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    n, seed, base, sims, cpuct = 6, 21, 5000, 20, 1.5
    fn = lambda b: mcts_oracle.hashed_evaluator(b, 11)
    plies = 16
    traces = [mcts_oracle.self_play_game(fn, sims, cpuct, seed, base + g, temp_threshold=6, max_plies=plies)[0]
              for g in range(n)]
    env = _engine(n, seed, base)
    ev = HostEvaluator(fn, n)
    mcts = BatchedMCTS(env, sims, cpuct, evaluator=ev, temp_threshold=6, want_leaf_states=True)
    for ply in range(plies):
        mcts.search()
        mcts.check_errors()
        counts, _ = mcts.root_counts()
        c = counts.cpu().numpy()
        acts = mcts.pick_actions().cpu().numpy()
        for g in range(n):
            got = {int(a): int(c[g][a]) for a in np.flatnonzero(c[g])}
            assert got == traces[g][ply]["counts"], (g, ply)
            assert int(acts[g]) == traces[g][ply]["action"]
        env.next_state(mcts.picked)


def test_uniform_full_games_vs_oracle():
    """BASELINE.json configs[2] semantics (uniform prior, 25 sims) on a small batch, every ply."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    n, seed, base = 3, 2, 300
    traces = [mcts_oracle.self_play_game(mcts_oracle.uniform_evaluator, 25, 1.5, seed, base + g)[0] for g in range(n)]
    env = _engine(n, seed, base)
    mcts = BatchedMCTS(env, 25, 1.5)
    for ply in range(48):
        mcts.play_ply()
        c = mcts.counts.cpu().numpy()
        for g in range(n):
            got = {int(a): int(c[g][a]) for a in np.flatnonzero(c[g])}
            assert got == traces[g][ply]["counts"], (g, ply)
            assert int(mcts.picked[g].item()) == traces[g][ply]["action"]
    mcts.check_errors()
    assert bool((env.game_ended() != 0).all())


def test_dropin_mcts_class_matches_reference_trace():
    """The reference-shaped MCTS(game, nnet, args).getActionProb surface on the 'hashed_s30' golden."""
    from nypc_yacht_auction_b200.mcts import MCTS
    from nypc_yacht_auction_b200.engine import BatchedYacht
    case = [c for c in load_cases() if c["name"] == "hashed_s30"][0]

    class Args(dict):
        __getattr__ = dict.__getitem__

    class Net:
        def predict(self, board):
            return mcts_oracle.hashed_evaluator(to_oracle_board(board))

    env = _engine(1, case["seed"], case["game"])
    args = Args(numMCTSSims=case["sims"], cpuct=case["cpuct"], search_seed=case["seed"], tree_id=case["game"], search_dice="philox")
    mcts = MCTS(None, Net(), args)
    from nypc_yacht_auction_b200.layout import planes_to_boards
    for ref in case["trace"][:20]:
        canon = planes_to_boards(env.canonical().cpu().numpy().view(np.uint32))[0]
        probs = mcts.getActionProb(canon, temp=1)
        total = sum(ref["counts"].values())
        assert len(probs) == 3226 and abs(sum(probs) - 1.0) < 1e-12
        for a, cnt in ref["counts"].items():
            assert probs[int(a)] == cnt / float(total)
        assert sum(1 for p in probs if p) == len(ref["counts"])
        env.next_state(torch.tensor([ref["action"]], dtype=torch.int32))


def test_pool_overflow_is_reported():
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    from nypc_yacht_auction_b200._lib import YachtB200Error
    env = _engine(2, 0, 0)
    mcts = BatchedMCTS(env, 25, 1.5, max_nodes=8)
    mcts.search()
    with pytest.raises(YachtB200Error):
        mcts.check_errors()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_fused_logits_expand_matches_softmax_mask_renorm(dtype):
    """ya_mcts_expand_logits: prior rows written from raw 16-bit logits (IEEE half or bfloat16) equal softmax -> mask
    -> renormalise (MCTS.py:86-91) computed in numpy, to float32 rounding (tolerance 2e-6 relative: exp differs by ulps)."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS

    class LogitEval:
        uniform = False
        returns_logits = True

        def __init__(self, n):
            g = torch.Generator().manual_seed(5)
            self.logits = (torch.randn((n, 3232), generator=g) * 3).to(dtype).cuda()
            self.v = torch.linspace(-0.9, 0.9, n).cuda()

        def __call__(self, features, need_eval, leaf_states):
            return self.logits, self.v

    n = 9
    legal_counts = set()
    for plies in (0, 5, 46, 47):                           # bid row (202), ten-dice score rows, five-dice rows of round 13
        legal_counts |= _check_expand_rows(n, plies, LogitEval)
    assert 202 in legal_counts and max(legal_counts) >= 2016 and min(legal_counts) <= 12, legal_counts


def logit_row_slots(desc, n_legal):
    """Slot of every legal move's logit inside a logit-row node's logit area (csrc/ya_mcts.cu "Logit area"), and the
    area's size in words: ten-dice rows keep one 272-slot block per open category, the category's 252 logits starting
    (10 + 12 c) mod 16 slots into it; bid and five-dice rows are compact."""
    if desc >> 13:
        cats = [c for c in range(12) if (desc >> (1 + c)) & 1]
        slots = np.concatenate([272 * r + ((10 + 12 * c) & 15) + np.arange(252) for r, c in enumerate(cats)])
        assert len(slots) == n_legal
        return slots, 136 * len(cats)
    return np.arange(n_legal), (((n_legal + 1) // 2) + 3) & ~3


def decode_logit_row(pool, g, node, n_legal):
    """(priors float64[L], legal logits float64[L], group-maximum words, visited words) of a logit-row node (csrc/ya_mcts.cu
    ROWS_L16): P[k] = 2^(l[k] * log2(e) + off) from the row's 16-bit logits and the node's exponent offset."""
    nodes = pool.nodes[g, node].cpu().numpy().view(np.uint32)
    assert int(nodes[13]) == 2, "not a logit-row node"
    off = int(nodes[10])
    slots, lw = logit_row_slots(int(nodes[8]), n_legal)
    nb = (n_legal + 31) // 32
    words = pool.arena[g, off:off + lw + 2 * nb].cpu().numpy().view(np.uint32)
    halves = words[:lw].view(np.uint16)[slots]
    if pool.rows == 2:
        lg = halves.view(np.float16).astype(np.float64)
    else:
        lg = (halves.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    off_p = float(nodes[14:15].view(np.float32)[0])
    pri = np.exp2(lg * float(np.float32(1.4426950408889634)) + off_p)
    gm = words[lw:lw + nb]
    return pri, lg, gm, words[lw + nb:lw + 2 * nb]


def _check_expand_rows(n, plies, LogitEval):
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    env = _engine(n, 3, 40)
    for ply in range(plies):
        env.play_ply(masks=None, auto_reset=False)
    ev = LogitEval(n)
    mcts = BatchedMCTS(env, 4, 1.5, evaluator=ev)
    assert mcts.rows in (2, 3) and not mcts.scatter        # dense 16-bit logits in, logit rows in the pool
    mcts.simulate(0)                                       # expands every root (node 0 of every tree)
    mcts.check_errors()
    masks = env.valid_moves(states=env.canonical(), players=torch.ones(n, dtype=torch.int8, device="cuda")).cpu().numpy()
    lg = ev.logits.float().cpu().numpy()[:, :3226]
    counts = set()
    for g in range(n):
        legal = np.flatnonzero(masks[g])
        row, row_logits, gm, seen = decode_logit_row(mcts.pool, g, 0, len(legal))
        assert (row_logits == lg[g][legal]).all(), (plies, g)                     # the legal logits, compacted, untouched
        e = np.exp((lg[g] - lg[g].max()).astype(np.float32)).astype(np.float32)
        pi = e / e.sum(dtype=np.float32)
        p = pi * masks[g]
        p = p / np.sum(p)
        assert np.allclose(row, p[legal], rtol=1e-5, atol=1e-12), (plies, g)
        assert abs(float(row.sum()) - 1.0) < 1e-5
        # group maxima: 1 + bits of the largest prior of each 32 (nothing visited yet)
        got = (gm - 1).view(np.float32).astype(np.float64)
        want = np.array([row[k:k + 32].max() for k in range(0, len(legal), 32)])
        assert (gm > 0).all() and np.allclose(got, want, rtol=1e-5), (plies, g)
        assert not seen.any()
        counts.add(len(legal))
    return counts


def _perturbed_net(seed, **kw):
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    torch.manual_seed(seed)
    net = YachtPolicyValueNet(**kw).cuda().eval()
    with torch.no_grad():
        for p in net.parameters():                     # non-trivial LayerNorm affine parameters and biases
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
    return net


PRECISIONS = [("fp16", torch.float16), ("bf16", torch.bfloat16)]


@pytest.mark.parametrize("precision,dtype", PRECISIONS, ids=["fp16", "bf16"])
def test_fused_evaluator_matches_module_forward(precision, dtype):
    """FusedYachtEvaluator (one tcgen05 kernel, 16-bit operands, float32 accumulation) against the fp32 module
    forward.  The operand format (ulp 2^-11 for fp16, 2^-8 for bf16) through 13 layers is the error floor, so the
    yardstick is PyTorch's own forward of the same module in that dtype: the kernel must be at least as close to fp32
    (stated tolerance: 1.25x the torch 16-bit forward's max error + 0.01), and the policy within 2 % total variation
    (fp16: 0.5 %)."""
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    from torch_evaluator import TorchEvaluator
    net = _perturbed_net(0)
    env = _engine(300, 1, 1)
    for _ in range(7):
        env.play_ply(masks=None, auto_reset=False)
    x = env.features()
    logits, v = FusedYachtEvaluator(net, max_batch=512, precision=precision)(x)
    t_logits, t_v = TorchEvaluator(net, dtype=dtype, fused_logits=True)(x)
    with torch.no_grad():
        ref_logits, ref_v = net(x)
    assert logits.shape == (300, 3232) and logits.dtype == dtype
    err = (logits[:, :3226].float() - ref_logits).abs().max().item()
    err_torch = (t_logits[:, :3226].float() - ref_logits).abs().max().item()
    assert err <= 1.25 * err_torch + 0.01, (err, err_torch)
    verr = (v - ref_v.reshape(-1)).abs().max().item()
    verr_torch = (t_v - ref_v.reshape(-1)).abs().max().item()
    assert verr <= 1.25 * verr_torch + 0.01, (verr, verr_torch)
    tv = 0.5 * (torch.softmax(logits[:, :3226].float(), 1) - torch.softmax(ref_logits, 1)).abs().sum(1).max().item()
    assert tv < (0.005 if precision == "fp16" else 0.02), tv


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_two_tile_schedule_is_bit_identical(precision):
    """ya_nn_forward_tiles: the two-tiles-per-CTA kernel (tensor core on one tile under the other tile's epilogue) computes
    every row with the same arithmetic as the one-tile kernel -- dense logits, values and row maxima are bit-identical, for
    full and ragged waves (rows beyond n, a last CTA pair with one to four tiles), so the choice of schedule (by wave size)
    cannot break batch / shard invariance."""
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    net = _perturbed_net(5)
    g = torch.Generator(device="cuda").manual_seed(11)
    for n in (1, 129, 300, 513, 1100):
        x = torch.rand((n, 59), device="cuda", generator=g)
        outs = []
        for tiles in (1, 2):
            ev = FusedYachtEvaluator(net, n, precision=precision, tiles_per_cta=tiles)
            logits, v = ev(x)
            outs.append((logits[:, :3226].clone(), v.clone(), ev.last_row_max.clone()))
        # a feature matrix that is only 4-byte aligned takes the per-thread loads instead of the bulk copy of whole tiles
        x4 = torch.empty(n * 59 + 1, device="cuda")[1:].view(n, 59)
        x4.copy_(x)
        assert x4.data_ptr() % 16 != 0
        ev = FusedYachtEvaluator(net, n, precision=precision, tiles_per_cta=1)
        logits, v = ev(x4)
        outs.append((logits[:, :3226].clone(), v.clone(), ev.last_row_max.clone()))
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a.view(torch.uint8) if a.dtype == torch.bfloat16 else a, b.view(torch.uint8) if b.dtype == torch.bfloat16 else b), n
    with pytest.raises(ValueError):
        FusedYachtEvaluator(net, 8, tiles_per_cta=3)


@pytest.mark.parametrize("precision,tiles", [("fp16", 1), ("bf16", 1), ("fp16", 2)])
def test_scattered_rows_equal_dense_logit_rows(precision, tiles):
    """the policy-head epilogue writes each leaf's legal logits straight into its row of the tree pool
    (ya_nn_forward with scatter targets).  Against the same kernel writing the dense [n, 3232] matrix and
    ya_mcts_expand_logits compacting it: identical rows, visit counts, Q values (bit patterns) and moves, over bid plies,
    ten-dice score plies (every category pattern the games reach) and -- via late plies -- five-dice rows."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
    net = _perturbed_net(21)

    class Dense:                                           # same forward kernel, no scatter capability
        uniform = False
        returns_logits = True

        def __init__(self, ev):
            self.ev, self.op_dtype, self.last_row_max = ev, ev.op_dtype, None

        def __call__(self, features, need_eval, leaf_states):
            out = self.ev(features)
            self.last_row_max = self.ev.last_row_max
            return out

    n, sims = 160, 20
    for start in (0, 40):                                  # from the deal, and from ply 40 (rounds 11-13: few open categories, 5 dice)
        runs = []
        for make in (lambda: FusedYachtEvaluator(net, n, precision=precision, tiles_per_cta=tiles),
                     lambda: Dense(FusedYachtEvaluator(net, n, precision=precision, tiles_per_cta=tiles))):
            env = _engine(n, 8, 300)
            for _ in range(start):
                env.play_ply(masks=None, auto_reset=False)
            m = BatchedMCTS(env, sims, 1.5, evaluator=make())
            trace = []
            for ply in range(8):
                m.search()
                c, v, q, k = m.root_counts(with_q=True)
                trace.append((c.clone(), v.clone(), q.clone(), k.clone(), m.pick_actions().clone()))
                env.next_state(m.picked)
            m.check_errors()
            runs.append((m.scatter, trace, m.pool.meta[:, :2].clone()))
        assert runs[0][0] and not runs[1][0]
        for a, b in zip(runs[0][1], runs[1][1]):
            for x, y in zip(a, b):
                assert torch.equal(x, y)
        assert torch.equal(runs[0][2], runs[1][2])         # same node counts and arena tops


def test_grouped_streams_give_identical_trees():
    """Splitting the games into two groups on two streams (overlap of tree kernels and forward) must not
    change anything: same counts and actions as the single-group run, eager and under CUDA-graph replay."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    torch.manual_seed(1)
    net = YachtPolicyValueNet().cuda().eval()
    n, sims = 96, 12
    results = []
    for groups, graph in ((1, False), (2, False), (3, True)):
        env = _engine(n, 5, 900)
        mcts = BatchedMCTS(env, sims, 1.5, evaluator=FusedYachtEvaluator(net, n), groups=groups)
        if graph:
            mcts.capture_graph()
        trace = []
        for ply in range(7):
            mcts.search()
            counts, visits = mcts.root_counts()
            trace.append((counts.cpu().clone(), visits.cpu().clone(), mcts.pick_actions().cpu().clone()))
            env.next_state(mcts.picked)
        mcts.check_errors()
        results.append(trace)
    for other in results[1:]:
        for (c0, v0, a0), (c1, v1, a1) in zip(results[0], other):
            assert torch.equal(c0, c1) and torch.equal(v0, v1) and torch.equal(a0, a1)


def test_batched_arena_vs_oracle():
    """BatchedArena (two searchers per game, greedy moves) against two oracle TreeSearch objects playing the
    same games with the same Philox streams."""
    from nypc_yacht_auction_b200.coach import BatchedArena
    from nypc_yacht_auction_b200.mcts import UniformEvaluator
    from oracle import philox
    n, sims, seed, base = 3, 8, 13, 70

    def oracle_game(gid, first_is_a):
        trees = {1: mcts_oracle.TreeSearch(mcts_oracle.uniform_evaluator, sims, 1.5, seed, gid),
                 -1: mcts_oracle.TreeSearch(mcts_oracle.uniform_evaluator, sims, 1.5, seed, gid)}
        board = yr.new_game(philox.Draw(seed, gid, 0, 0, philox.TAG_INIT))
        cur, ply = 1, 0
        while yr.outcome(board, cur) == 0:
            counts = trees[cur].root_counts(yr.canonical(board, cur), ply)
            best = np.flatnonzero(counts == counts.max())
            word = philox.draw_words(seed, gid, 0, ply, philox.TAG_ACTION)[3]
            a = int(best[(word * len(best)) >> 32])
            board, cur = yr.next_state(board, cur, a, philox.Draw(seed, gid, 0, ply, philox.TAG_REAL))
            ply += 1
        return yr.outcome(board, 1)

    arena = BatchedArena(n, sims, UniformEvaluator(), UniformEvaluator(), seed=seed, game_base=base)
    a, b, d = arena.play_games()
    exp_a = exp_b = exp_d = 0
    for e in range(2):
        for g in range(n):
            r = oracle_game(base + e * n + g, e == 0)
            if abs(r) < 0.5:
                exp_d += 1
            elif (r > 0) == (e == 0):
                exp_a += 1
            else:
                exp_b += 1
    assert (a, b, d) == (exp_a, exp_b, exp_d) and a + b + d == 2 * n


def test_batched_arena_against_scripted_greedy_seat():
    """An MCTS population against the device GreedyYachtPlayer (seat "greedy"), both seatings, against the oracle's
    tree search and the oracle's restatement of the heuristic on the same Philox streams."""
    from nypc_yacht_auction_b200.coach import BatchedArena
    from nypc_yacht_auction_b200.mcts import UniformEvaluator
    from oracle import greedy_oracle, philox
    n, sims, seed, base = 4, 8, 17, 300

    def oracle_game(gid, mcts_seat):
        tree = mcts_oracle.TreeSearch(mcts_oracle.uniform_evaluator, sims, 1.5, seed, gid)
        board = yr.new_game(philox.Draw(seed, gid, 0, 0, philox.TAG_INIT))
        cur, ply = 1, 0
        while yr.outcome(board, cur) == 0:
            canon = yr.canonical(board, cur)
            if cur == mcts_seat:
                counts = tree.root_counts(canon, ply)
                best = np.flatnonzero(counts == counts.max())
                word = philox.draw_words(seed, gid, 0, ply, philox.TAG_ACTION)[3]
                a = int(best[(word * len(best)) >> 32])
            else:
                a, legal = greedy_oracle.greedy_action(canon)
                assert legal                                           # the overflow fallback (Q11) has its own test
            board, cur = yr.next_state(board, cur, a, philox.Draw(seed, gid, 0, ply, philox.TAG_REAL))
            ply += 1
        return yr.outcome(board, 1)

    arena = BatchedArena(n, sims, UniformEvaluator(), "greedy", seed=seed, game_base=base)
    a, b, d = arena.play_games()
    exp = [0, 0, 0]
    for e in range(2):
        for g in range(n):
            r = oracle_game(base + e * n + g, 1 if e == 0 else -1)
            exp[2 if abs(r) < 0.5 else (0 if (r > 0) == (e == 0) else 1)] += 1
    assert [a, b, d] == exp and a + b + d == 2 * n


@pytest.mark.parametrize("precision,dtype", PRECISIONS, ids=["fp16", "bf16"])
def test_whole_forward_kernel(precision, dtype):
    """csrc/ya_forward.cu: features -> logits / values in one tcgen05 kernel, against PyTorch's forward of the same module
    in the same operand dtype and the fp32 module, on a ragged row count (not a multiple of the 128-row tile); and batch
    invariance: a row evaluated alone, in a small batch or in a large one gives bit-identical logits and value."""
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    from torch_evaluator import TorchEvaluator
    net = _perturbed_net(3)
    n = 777
    env = _engine(n, 6, 6)
    for _ in range(11):
        env.play_ply(masks=None, auto_reset=False)
    x = env.features()
    one = FusedYachtEvaluator(net, max_batch=n, precision=precision)
    lf, vf = one(x)
    lf, vf = lf.clone(), vf.clone()
    ls, vs = TorchEvaluator(net, dtype=dtype, fused_logits=True)(x)
    with torch.no_grad():
        ref_logits, ref_v = net(x)
    err_fast = (lf[:, :3226].float() - ref_logits).abs().max().item()
    err_slow = (ls[:, :3226].float() - ref_logits).abs().max().item()
    assert err_fast <= 1.25 * err_slow + 0.01, (err_fast, err_slow)
    v_fast = (vf - ref_v.reshape(-1)).abs().max().item()
    v_slow = (vs - ref_v.reshape(-1)).abs().max().item()
    assert v_fast <= 1.25 * v_slow + 0.01, (v_fast, v_slow)
    tv = 0.5 * (torch.softmax(lf[:, :3226].float(), 1) - torch.softmax(ref_logits, 1)).abs().sum(1).max().item()
    assert tv < (0.005 if precision == "fp16" else 0.02), tv
    assert torch.equal(one.last_row_max, lf[:, :3226].float().max(dim=1).values)     # epilogue by-product for expand
    # batch invariance (bitwise)
    for lo, hi in ((0, 1), (5, 6), (100, 229), (640, 777)):
        l2, v2 = one(x[lo:hi].contiguous())
        assert torch.equal(l2[:, :3226], lf[lo:hi, :3226]) and torch.equal(v2, vf[lo:hi])


def test_evaluator_abi_rejects_bad_arguments():
    """The C entry points of the leaf evaluator report misuse as CUDA error codes instead of launching."""
    from nypc_yacht_auction_b200 import _lib
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    lib = _lib.load()
    ev = FusedYachtEvaluator(YachtPolicyValueNet().cuda().eval(), 256)
    x = torch.zeros((256, 59), device="cuda")
    ev(x)                                                                            # allocates the dense logit matrix
    s = _lib.current_stream()
    args = lambda logits_ptr, n, dst=None, desc=None: (
        _lib.ptr(x), logits_ptr, _lib.ptr(ev.values), _lib.ptr(ev.row_max), _lib.ptr(ev.fw_w), _lib.ptr(ev.fw_p), ev.fw_off,
        ev.nblocks, n, ev.eps, 1, _lib.ptr(dst), _lib.ptr(desc), s)
    assert lib.ya_nn_forward(*args(_lib.ptr(ev.logits), 0)) == 0                     # nothing to do
    assert lib.ya_nn_forward(*args(ev.logits.data_ptr() + 2, 256)) != 0              # logits not 32-byte aligned
    assert lib.ya_nn_forward(*args(None, 256)) != 0                                  # neither a dense matrix nor scatter targets
    dst = torch.zeros(256, dtype=torch.int64, device="cuda")
    assert lib.ya_nn_forward(*args(None, 256, dst, None)) != 0                       # scatter targets come in pairs
    env = _engine(4, 1, 1)
    from nypc_yacht_auction_b200.mcts import BatchedMCTS
    m = BatchedMCTS(env, 2, 1.5, evaluator=ev.with_private_buffers(4))
    v = torch.zeros(4, device="cuda")
    bad_ld = lib.ya_mcts_expand_logits(m.pool.ref, _lib.ptr(ev.logits), 1, 3226, None, _lib.ptr(v), None, _lib.ptr(m.err_flag), s)
    assert bad_ld != 0                                                               # row stride must be >= 3232 and % 8 == 0
    torch.cuda.synchronize()


def test_full_size_mcts_properties():
    """BASELINE.json configs[2] and [3] at their full sizes (4,096 games x 25 sims uniform; 16,384 games x 100 sims
    with the network), checked through size-independent properties of MCTS.getActionProb (MCTS.py:28-54): the
    root's visit counts sum to Ns, a fresh root has exactly numMCTSSims - 1 child visits, every visited and every
    picked action is legal, no pool overflow, and the first games of the big batch equal a small batch's games
    (shard invariance)."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
    from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
    torch.manual_seed(4)
    net = YachtPolicyValueNet().cuda().eval()
    for n, sims, make_ev in ((4096, 25, lambda m: None), (16384, 100, lambda m: FusedYachtEvaluator(net, m))):
        env = _engine(n, 8, 70000)
        small = _engine(64, 8, 70000)
        mcts = BatchedMCTS(env, sims, 1.5, evaluator=make_ev(n))
        ref = BatchedMCTS(small, sims, 1.5, evaluator=make_ev(64))
        for ply in range(3):
            masks = env.valid_moves(states=env.canonical(), players=torch.ones(n, dtype=torch.int8, device="cuda")).bool()
            mcts.search()
            ref.search()
            counts, visits = mcts.root_counts()
            rcounts, _ = ref.root_counts()
            assert torch.equal(counts[:64], rcounts)
            assert bool((counts.sum(1) == visits).all())                           # sum of Nsa over the root's edges = Ns
            if ply == 0:
                assert bool((counts.sum(1) == sims - 1).all())                     # fresh root: the first search only expands
            assert not bool((counts.bool() & ~masks).any())                        # visits only on legal moves
            picked = mcts.pick_actions()
            assert bool(masks.gather(1, picked.long().unsqueeze(1)).all())
            assert torch.equal(picked[:64], ref.pick_actions())
            env.next_state(picked)
            small.next_state(ref.picked)
        mcts.check_errors()
        ref.check_errors()
        assert int(mcts.pool.node_counts().max().item()) <= mcts.pool.max_nodes
        del mcts, ref, env, small
        torch.cuda.empty_cache()


def test_expand_falls_back_to_uniform_when_every_legal_move_underflows():
    """MCTS.py:92-101: if the masked policy sums to 0 (here: softmax underflows on every legal move because an
    illegal move carries all the mass), the reference falls back to Ps = valids / sum(valids)."""
    from nypc_yacht_auction_b200.mcts import BatchedMCTS

    class Underflow:
        uniform = False
        returns_logits = True

        def __init__(self, n):
            self.logits = torch.zeros((n, 3232), dtype=torch.bfloat16, device="cuda")
            self.logits[:, :202] = -200.0                    # every bid far below the (illegal) score moves
            self.v = torch.zeros(n, device="cuda")

        def __call__(self, features, need_eval, leaf_states):
            return self.logits, self.v

    n = 5
    env = _engine(n, 2, 9)                                   # ply 0: bid rows, 202 legal moves
    mcts = BatchedMCTS(env, 4, 1.5, evaluator=Underflow(n))
    mcts.simulate(0)
    mcts.check_errors()
    nodes = mcts.pool.nodes.cpu().numpy().view(np.uint32)
    for g in range(n):                                       # the leaf became a constant-prior node: P = 1 / 202 for every bid
        assert int(nodes[g, 0, 13]) == 1
        assert nodes[g, 0, 14:15].view(np.float32)[0] == np.float32(1.0) / np.float32(202)
    for sim in range(1, 4):                                  # and the search walks it like one: lowest unvisited bid first
        mcts.simulate(sim)
    mcts.check_errors()
    counts, visits = mcts.root_counts()
    assert counts[:, :3].tolist() == [[1, 1, 1]] * n and int(counts.sum()) == 3 * n and visits.tolist() == [3] * n


@pytest.mark.parametrize("nblocks,precision", [(1, "fp16"), (3, "bf16")])
def test_whole_forward_kernel_other_depths(nblocks, precision):
    """The forward kernel loops over the residual blocks it is given: shallower trunks against the float32 module."""
    from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
    net = _perturbed_net(10 + nblocks, nblocks=nblocks)
    env = _engine(300, 4, 4)
    for _ in range(9):
        env.play_ply(masks=None, auto_reset=False)
    x = env.features()
    ev = FusedYachtEvaluator(net, 300, precision=precision)
    assert ev.nblocks == nblocks
    logits, values = ev(x)
    with torch.no_grad():
        ref_logits, ref_v = net(x)
    tv = 0.5 * (torch.softmax(logits[:, :3226].float(), 1) - torch.softmax(ref_logits, 1)).abs().sum(1).max().item()
    assert (logits[:, :3226].float() - ref_logits).abs().max().item() < 0.15 and tv < 0.02
    assert (values - ref_v.reshape(-1)).abs().max().item() < 0.05
