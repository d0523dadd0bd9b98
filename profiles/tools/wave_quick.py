"""select / forward / expand per network wave at a score ply and a bid ply, the forward alone on the last wave's real
leaves (back to back, and with an L2-sized memset in between), plus one full episode rate."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200 import mcts_bench, _lib
from nypc_yacht_auction_b200.engine import BatchedYacht
from nypc_yacht_auction_b200.mcts import BatchedMCTS, FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
dev = torch.device("cuda", 0)
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(pos[0]) if pos else 16384
episode = "--episode" in sys.argv
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev).eval()
tiles = next((int(a.split('=')[1]) for a in sys.argv if a.startswith('--tiles=')), 0)
ev0 = FusedYachtEvaluator(net, n, tiles_per_cta=tiles)
def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plies = next(([int(x) for x in a.split('=')[1].split(',')] for a in sys.argv if a.startswith('--plies=')), [6, 8])
for ply in plies:
    env = BatchedYacht(n, seed=0, game_base=0, device=dev)
    m = BatchedMCTS(env, 100, 1.5, evaluator=ev0)
    for _ in range(ply):
        m.play_ply()
    grp, lib, s = m.groups[0], m.lib, _lib.current_stream()
    ev = grp.evaluator
    waves = 60
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(waves)]
    def select(sim):
        _lib.check(lib.ya_mcts_select(grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
                                      env.seed, env.game_base, sim, None, None, m.cpuct, None, _lib.ptr(grp.features),
                                      _lib.ptr(grp.need_eval), None, m.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc),
                                      _lib.ptr(m.err_flag), s), "ya_mcts_select")
    def expand(v):
        _lib.check(lib.ya_mcts_expand_logits(grp.ref, None, 1 if m.rows == 2 else 0, 0, _lib.ptr(ev.last_row_max), _lib.ptr(v), None,
                                             _lib.ptr(m.err_flag), s), "ya_mcts_expand_logits")
    for sim in range(waves):
        a = marks[sim]
        a[0].record(); select(sim); a[1].record()
        pi, v = ev(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
        a[2].record(); expand(v); a[3].record()
    torch.cuda.synchronize(dev)
    tot = [sum(a[i].elapsed_time(a[i + 1]) for a in marks[5:]) * 1e3 / (waves - 5) for i in range(3)]
    # the forward alone on the leaves of one more descent (rows are rewritten with the same logits: harmless)
    select(waves)
    desc = grp.leaf_desc
    kinds = ((desc & 1) != 0).sum().item(), ((desc >> 13) != 0).sum().item(), (grp.leaf_dst == 0).sum().item()
    e0, e1 = ev_pair(); e0.record()
    for _ in range(20):
        pi, v = ev(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
    e1.record(); torch.cuda.synchronize(dev)
    alone = e0.elapsed_time(e1) * 50
    junk = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(10):
        junk.fill_(1)
        e0, e1 = ev_pair(); e0.record()
        pi, v = ev(grp.features, grp.need_eval, None, scatter=(grp.leaf_dst, grp.leaf_desc))
        e1.record(); torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1) * 1e3)
    expand(v)
    torch.cuda.synchronize(dev)
    if "--timeline" in sys.argv:      # CTA 0's stage timeline on these real leaves (profiling build of ya_forward.cu, see forward_timeline.py)
        import ctypes
        dbg = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dbg", "libfwd_tl.so"))
        vp = ctypes.c_void_p
        dbg.ya_nn_forward_tiles.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_int, vp, vp, ctypes.c_int, vp]
        for _ in range(3):
            assert dbg.ya_nn_forward_tiles(grp.features.data_ptr(), None, ev.values.data_ptr(), ev.row_max.data_ptr(), ev.fw_w.data_ptr(), ev.fw_p.data_ptr(),
                                           ev.fw_off, ev.nblocks, n, ev.eps, 1, grp.leaf_dst.data_ptr(), grp.leaf_desc.data_ptr(), 1, s) == 0
            torch.cuda.synchronize(dev)
        buf = (ctypes.c_ulonglong * 1024)()
        assert dbg.ya_debug_forward_timeline(buf) == 0
        t = list(buf)
        base = 3 * (2 + 2 * ev.nblocks)
        pol = t[base:base + 27]
        print("  timeline (real leaves): trunk done at %.2f us, first policy accumulator at %.2f us, per tile: %s | total %.2f us" % (
            (t[3 * (1 + 2 * ev.nblocks)] - t[0]) / 1e3, (pol[0] - t[0]) / 1e3, " ".join("%.2f" % ((pol[i + 1] - pol[i]) / 1e3) for i in range(26)), (pol[26] - t[0]) / 1e3))
        cb = (ctypes.c_ulonglong * 4096)()
        assert dbg.ya_debug_forward_cta_times(cb) == 0
        c = list(cb)
        nb = 2 * ((n + 255) // 256)
        t0 = min(c[4 * i] for i in range(nb))
        st = sorted((c[4 * i] - t0) / 1e3 for i in range(nb))
        tr = sorted((c[4 * i + 1] - c[4 * i]) / 1e3 for i in range(nb))
        en = sorted((c[4 * i + 2] - t0) / 1e3 for i in range(nb))
        du = sorted((c[4 * i + 2] - c[4 * i]) / 1e3 for i in range(nb))
        q = lambda a: "min %.1f  median %.1f  p90 %.1f  max %.1f" % (a[0], a[len(a) // 2], a[int(len(a) * 0.9)], a[-1])
        print("  per CTA (us): start after first: %s | trunk: %s | duration: %s | end: %s | SMs used %d" % (q(st), q(tr), q(du), q(en), len({c[4 * i + 3] for i in range(nb)})))
    print("n=%d ply %d: select %.1f  forward %.1f  expand %.1f us  (wave %.1f us = %.3g sims/s); forward alone %.1f us, after a 160 MB memset %.1f us; leaves: %d bid rows, %d ten-dice rows, %d not evaluated"
          % (n, ply, tot[0], tot[1], tot[2], sum(tot), n / sum(tot) * 1e6, alone, sorted(ts)[len(ts) // 2], kinds[0], kinds[1], kinds[2]))
    del m, env, junk
    torch.cuda.empty_cache()
if episode:
    out = mcts_bench.selfplay_block(torch, dev, None, 0, 1, n, 100, ev0, 0, 1, 1)
    print("episode: %.1f ms  %.4g sims/s" % (out["ms_per_step"], out["sims_per_sec"]))
