"""Feasibility probe: SM partitioning with CUDA green contexts.  The forward kernel (one 128-leaf CTA per SM, tensor-bound) on
F SMs, the latency-bound tree kernels (select, expand) on the remaining SMs -- do they keep their speed, and do they overlap?
    python profiles/tools/partition_probe.py [F ...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cuda.bindings import driver as drv
from nypc_yacht_auction_b200 import _lib
from nypc_yacht_auction_b200.engine import BatchedYacht
from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet


def ck(res):
    err = res[0]
    assert int(err) == 0, err
    return res[1:] if len(res) > 2 else res[1]


dev_t = torch.device("cuda", 0)
torch.zeros(1, device=dev_t)                                   # primary context
dev = ck(drv.cuDeviceGet(0))
sm_all = ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("SMs:", sm_all.sm.smCount, flush=True)


def partition(fwd_sms):
    groups, n, rest = ck(drv.cuDevSmResourceSplitByCount(1, sm_all, 0, fwd_sms))
    out = []
    for r in (groups[0], rest):
        desc = ck(drv.cuDevResourceGenerateDesc([r], 1))
        g = ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = ck(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append((r.sm.smCount, torch.cuda.ExternalStream(int(s), device=dev_t)))
    return out


torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev_t)


def timed(stream, fn, reps):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for fwd_sms in [int(a) for a in sys.argv[1:]] or [96, 104, 112]:
    (nf, s_fwd), (nt, s_tree) = partition(fwd_sms)
    leaves = nf * 128
    ev = FusedYachtEvaluator(net, leaves)
    x = torch.rand((leaves, 59), device=dev_t)
    env = BatchedYacht(leaves, seed=2, device=dev_t)
    m = BatchedMCTS(env, 100, 1.5, evaluator=FusedYachtEvaluator(net, leaves))
    for _ in range(6):
        m.play_ply()
    torch.cuda.synchronize()
    grp, lib = m.groups[0], m.lib
    sim = [0]

    def tree_wave():                                           # select -> (skip forward: rows keep the last logits) -> expand
        s = _lib.current_stream()
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                      env.seed, env.game_base, sim[0], None, None, m.cpuct, None, _lib.ptr(grp.features),
                                      _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                      _lib.ptr(m.err_flag), s), "select")
        grp.evaluator(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
        e = grp.evaluator
        _lib.check(lib.ya_mcts_expand_logits(grp.ref, None, 1, 0, _lib.ptr(e.last_row_max), _lib.ptr(e.values[:env.n]), None,
                                             _lib.ptr(m.err_flag), s), "expand")
        sim[0] += 1

    def tree_only():
        s = _lib.current_stream()
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                      env.seed, env.game_base, sim[0], None, None, m.cpuct, None, _lib.ptr(grp.features),
                                      _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                      _lib.ptr(m.err_flag), s), "select")
        e = grp.evaluator
        _lib.check(lib.ya_mcts_expand_logits(grp.ref, None, 1, 0, _lib.ptr(e.last_row_max), _lib.ptr(e.values[:env.n]), None,
                                             _lib.ptr(m.err_flag), s), "expand")
        sim[0] += 1

    main = torch.cuda.current_stream()
    t_wave_full = timed(main, tree_wave, 20)                   # all SMs: select + forward + expand, serial
    sim[0] = 30
    t_fwd_full = timed(main, lambda: ev(x), 20)
    t_fwd_part = timed(s_fwd, lambda: ev(x), 20)
    # tree kernels alone (no fresh logits: the rows keep the previous wave's; timing only) on all SMs and on the small partition
    # (run a real wave in between so the trees keep growing like in a search)
    def tree_pair_time(stream):
        tot = 0.0
        for _ in range(10):
            with torch.cuda.stream(main):
                tree_wave()
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s = _lib.current_stream()
                _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                              env.seed, env.game_base, sim[0], None, None, m.cpuct, None, _lib.ptr(grp.features),
                                              _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                              _lib.ptr(m.err_flag), s), "select")
                e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
            with torch.cuda.stream(main):                      # finish that wave normally
                e = grp.evaluator
                e(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
                _lib.check(lib.ya_mcts_expand_logits(grp.ref, None, 1, 0, _lib.ptr(e.last_row_max), _lib.ptr(e.values[:env.n]), None,
                                                     _lib.ptr(m.err_flag), _lib.current_stream()), "expand")
            sim[0] += 1
            torch.cuda.synchronize()
        return tot * 1e3 / 10
    t_sel_full = tree_pair_time(main)
    t_sel_part = tree_pair_time(s_tree)
    # concurrency: forward on its partition while select runs on the other one
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(main):
        tree_wave()
    torch.cuda.synchronize()
    e0.record(main)
    s_fwd.wait_stream(main); s_tree.wait_stream(main)
    with torch.cuda.stream(s_fwd):
        ev(x)
    with torch.cuda.stream(s_tree):
        s = _lib.current_stream()
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                      env.seed, env.game_base, sim[0], None, None, m.cpuct, None, _lib.ptr(grp.features),
                                      _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                      _lib.ptr(m.err_flag), s), "select")
    main.wait_stream(s_fwd); main.wait_stream(s_tree)
    e1.record(main)
    torch.cuda.synchronize()
    m.check_errors()
    print("forward SMs %d / tree SMs %d, %d leaves: wave(all SMs) %.1f us | forward: all SMs %.1f, its partition %.1f | select: all SMs %.1f, "
          "its partition %.1f | forward || select on their partitions: %.1f us" % (nf, nt, leaves, t_wave_full, t_fwd_full, t_fwd_part, t_sel_full,
                                                                                    t_sel_part, e0.elapsed_time(e1) * 1e3), flush=True)
    del m, env, ev
    torch.cuda.empty_cache()
