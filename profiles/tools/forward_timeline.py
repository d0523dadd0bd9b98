"""Stage-by-stage timeline of CTA 0 of ya_k_forward (globaltimer stamps; profiling build with -DYA_FWD_TIMELINE)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from nypc_yacht_auction_b200 import _lib, build
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
extra = [a for a in sys.argv[1:] if a.startswith("-D")]
sys.argv = [a for a in sys.argv if not a.startswith("-D")]
dbg = os.path.join(ROOT, "profiles", "tools", "_dbg", "libfwd_tl%s.so" % "".join(extra).replace("-D", "_"))
src = os.path.join(build.CSRC, "ya_forward.cu")
if not os.path.exists(dbg) or os.path.getmtime(dbg) < os.path.getmtime(src):
    os.makedirs(os.path.dirname(dbg), exist_ok=True)
    subprocess.check_call(["nvcc"] + build.NVCC_FLAGS + ["-DYA_FWD_TIMELINE"] + extra + ["-o", dbg, src])
if not torch.cuda.is_available():
    sys.exit(0)
lib = ctypes.CDLL(dbg)
dev = torch.device("cuda", 0)
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(pos[0]) if pos else 16384
ev = FusedYachtEvaluator(YachtPolicyValueNet().to(dev).eval(), n)
x = torch.rand((n, 59), device=dev)
vp = ctypes.c_void_p
lib.ya_nn_forward_tiles.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_int, vp, vp, ctypes.c_int, vp]
tiles = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--tiles=")), 0)
logits = torch.empty((n, 3232), dtype=torch.float16, device=dev)
# --scatter=bid|score|mix: the MCTS mode (legal logits into per-leaf rows, no dense matrix) with synthetic leaves: bid rows
# (208 slots) and / or ten-dice rows with every category open (12 x 272 slots)
mode = next((a.split("=")[1] for a in sys.argv if a.startswith("--scatter=")), None)
sc_dst = sc_desc = None
if mode:
    # --stride-mb=S: rows S MB apart, like the leaf rows of a per-game tree pool (one 2 MB page per leaf)
    stride_mb = next((float(a.split("=")[1]) for a in sys.argv if a.startswith("--stride-mb=")), None)
    row_elems = 12 * 272 + 16 if stride_mb is None else int(stride_mb * (1 << 20)) // 2
    area = torch.zeros((n, row_elems), dtype=torch.float16, device=dev)
    sc_dst = (area.data_ptr() + torch.arange(n, device=dev, dtype=torch.int64) * area.stride(0) * 2).contiguous()
    bid = torch.ones(n, dtype=torch.int32, device=dev)
    score = torch.full((n,), (1 << 13) | (0xFFF << 1), dtype=torch.int32, device=dev)
    mode = mode.split(",")[0]
    if mode == "score-hot":          # every 256-leaf cluster writes the same 256 rows: the stores stay in L2 (no HBM write-back)
        sc_dst = (area.data_ptr() + (torch.arange(n, device=dev, dtype=torch.int64) % 256) * area.stride(0) * 2).contiguous()
    sc_desc = {"bid": bid, "score": score, "score-hot": score, "mix": torch.where(torch.arange(n, device=dev) % 2 == 0, bid, score)}[mode].contiguous()
    logits = None
def run():
    return lib.ya_nn_forward_tiles(x.data_ptr(), logits.data_ptr() if logits is not None else None, ev.values.data_ptr(), ev.row_max.data_ptr(), ev.fw_w.data_ptr(),
                                   ev.fw_p.data_ptr(), ev.fw_off, ev.nblocks, n, ev.eps, 1, sc_dst.data_ptr() if mode else None,
                                   sc_desc.data_ptr() if mode else None, tiles, torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    rc = run()
    assert rc == 0, rc
    torch.cuda.synchronize()
two = tiles == 2 or (tiles == 0 and n > 148 * 128)
buf = (ctypes.c_ulonglong * 1024)()
assert lib.ya_debug_forward_timeline(buf) == 0
t = list(buf)
buf2 = (ctypes.c_ulonglong * 1024)()
assert lib.ya_debug_forward_timeline2(buf2) == 0
t2 = list(buf2)
if not two:
    nst = 2 + 2 * ev.nblocks                    # stages: input, trunk layers, value head
    names = ["input"] + ["trunk%d" % i for i in range(2 * ev.nblocks)] + ["value"]
    t0 = t[0]
    print("stage        start_us  wait_weights  mma   epilogue(until next stage's sync)")
    for k in range(nst):
        a, b, c = t[3 * k], t[3 * k + 1], t[3 * k + 2]
        nxt = t[3 * k + 3]
        print("%-10s %9.2f %9.2f %9.2f %9.2f" % (names[k], (a - t0) / 1e3, (b - a) / 1e3, (c - b) / 1e3, (nxt - c) / 1e3))
    base = 3 * nst
    pol = t[base:base + 27]
    print("policy: first accumulator ready at %.2f us; per tile (us):" % ((pol[0] - t0) / 1e3), " ".join("%.2f" % ((pol[i + 1] - pol[i]) / 1e3) for i in range(26)))
    print("total %.2f us" % ((pol[26] - t0) / 1e3))
    if t2[0]:
        print("trunk layer detail (us, thread 0): issue->h0_ready  pass1_h0  wait_h1  pass1_h1  stats  pass2 | total")
        for l in range(2 * ev.nblocks):
            q = t2[8 * l:8 * l + 8]
            d = [(q[i + 1] - q[i]) / 1e3 for i in range(6)]
            print("  L%-2d  %s | %.2f" % (l, "  ".join("%5.2f" % x for x in d), (q[6] - q[0]) / 1e3))
else:
    t0 = t2[1000]
    print("two tiles per CTA; epilogues of CTA 0 (us, thread 0): start | wait_h0  pass1_h0  wait_h1  pass1_h1  stats  pass2  close | total")
    for e in range(4 * ev.nblocks):
        q = t2[8 * e:8 * e + 8]
        d = [(q[i + 1] - q[i]) / 1e3 for i in range(7)]
        print("  L%-2d %s  %7.2f | %s | %.2f" % (e // 2, "XY"[e % 2], (q[0] - t0) / 1e3, "  ".join("%5.2f" % x for x in d), (q[7] - q[0]) / 1e3))
    names = ["features", "input weights", "X input MMA", "X input epilogue", "Y input MMA", "Y input epilogue", "(trunk) head parameters", "X heads", "Y heads", "X value", "Y value"]
    print("other stages (us since start, duration): " + "; ".join("%s %.2f (+%.2f)" % (names[i], (t[i + 1] - t0) / 1e3, (t[i + 1] - t[i]) / 1e3) for i in range(11)))
    t = t[12:]
    pol = t[0:53]
    print("policy: first accumulator ready at %.2f us; per (tile, X / Y) (us):" % ((pol[0] - t0) / 1e3), " ".join("%.2f" % ((pol[i + 1] - pol[i]) / 1e3) for i in range(52)))
    print("total %.2f us" % ((pol[52] - t0) / 1e3))
if t2[1002] > t2[1000]:
    print("SM clock inside the kernel: %.0f MHz (%d cycles in %.2f us)" % ((t2[1003] - t2[1001]) * 1e3 / (t2[1002] - t2[1000]),
          t2[1003] - t2[1001], (t2[1002] - t2[1000]) / 1e3))
cb = (ctypes.c_ulonglong * 4096)()
assert lib.ya_debug_forward_cta_times(cb) == 0
c = list(cb)
nb = 2 * ((n + (512 if two else 256) - 1) // (512 if two else 256))
tc0 = min(c[4 * i] for i in range(nb))
du = sorted((c[4 * i + 2] - c[4 * i]) / 1e3 for i in range(nb))
tr = sorted((c[4 * i + 1] - c[4 * i]) / 1e3 for i in range(nb))
print("per CTA (us): trunk done after min %.1f median %.1f max %.1f | duration min %.1f median %.1f max %.1f | last end %.1f | %d CTAs on %d SMs" % (
    tr[0], tr[len(tr) // 2], tr[-1], du[0], du[len(du) // 2], du[-1], max(c[4 * i + 2] - tc0 for i in range(nb)) / 1e3, nb, len({c[4 * i + 3] for i in range(nb)})))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
print("variant %s: %.1f us per forward (CUDA events, %d leaves, %s)" % ("".join(extra) or "production", e0.elapsed_time(e1) * 50.0, n,
      "scatter " + mode if mode else "dense logits"))
