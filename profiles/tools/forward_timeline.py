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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ev = FusedYachtEvaluator(YachtPolicyValueNet().to(dev).eval(), n)
x = torch.rand((n, 59), device=dev)
vp = ctypes.c_void_p
lib.ya_nn_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_int, vp, vp, vp]
logits = torch.empty((n, 3232), dtype=torch.float16, device=dev)
for _ in range(5):
    rc = lib.ya_nn_forward(x.data_ptr(), logits.data_ptr(), ev.values.data_ptr(), ev.row_max.data_ptr(), ev.fw_w.data_ptr(),
                           ev.fw_p.data_ptr(), ev.fw_off, ev.nblocks, n, ev.eps, 1, None, None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 1024)()
assert lib.ya_debug_forward_timeline(buf) == 0
t = list(buf)
nst = 2 + 2 * ev.nblocks                    # run_mma stages: input, trunk layers, value head
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.ya_nn_forward(x.data_ptr(), logits.data_ptr(), ev.values.data_ptr(), ev.row_max.data_ptr(), ev.fw_w.data_ptr(),
                      ev.fw_p.data_ptr(), ev.fw_off, ev.nblocks, n, ev.eps, 1, None, None, torch.cuda.current_stream().cuda_stream)
e1.record()
torch.cuda.synchronize()
print("variant %s: %.1f us per forward (CUDA events, %d leaves, dense logits)" % ("".join(extra) or "production", e0.elapsed_time(e1) * 50.0, n))
