"""CPU emulation of the forward kernel arithmetic (16-bit operands, float32 accumulation) with a float32 / fp16 / bf16 residual
stream, and of the reference CUDA predict (torch autocast fp16), against the float32 module: logit error and total variation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
torch.manual_seed(1234)
net = YachtPolicyValueNet().eval()
with torch.no_grad():
    for name, p in net.named_parameters():
        if name.endswith("bias") or "ln" in name or name in ("inp.1.weight", "pi_head.0.weight", "v_head.0.weight"):
            p.add_(0.1 * torch.randn_like(p))
sd = net.state_dict()
def r16(t, dt): return t.to(dt).float()
def lin(x, w, b, dt): return r16(x, dt) @ r16(w, dt).t() + b
def ln(x, g, b): return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, 1e-5)
silu = torch.nn.functional.silu
def fwd(x, dt, skipdt):
    h = silu(ln(lin(x, sd["inp.0.weight"], sd["inp.0.bias"], dt), sd["inp.1.weight"], sd["inp.1.bias"]))
    if skipdt is not None: h = r16(h, skipdt)
    for i in range(6):
        p = "blocks.%d." % i
        y = ln(silu(lin(h, sd[p+"fc1.weight"], sd[p+"fc1.bias"], dt)), sd[p+"ln1.weight"], sd[p+"ln1.bias"])
        y = ln(silu(lin(y, sd[p+"fc2.weight"], sd[p+"fc2.bias"], dt)), sd[p+"ln2.weight"], sd[p+"ln2.bias"])
        h = h + y
        if skipdt is not None: h = r16(h, skipdt)
    a = silu(ln(h, sd["pi_head.0.weight"], sd["pi_head.0.bias"]))
    pi = lin(a, sd["pi_head.2.weight"], sd["pi_head.2.bias"], dt)
    return pi, h
x = torch.rand(4096, 59)
with torch.no_grad():
    ref, href = net(x)[0], None
    for dt in (torch.float16, torch.bfloat16):
        for sk in (None, torch.float16, torch.bfloat16):
            pi, h = fwd(x, dt, sk)
            e = (pi - ref).abs()
            p0, p1 = torch.softmax(ref, -1), torch.softmax(pi, -1)
            tv = 0.5 * (p0 - p1).abs().sum(-1)
            print("operands %-8s skip %-8s: max|dlogit| %.4f  rms %.5f   TV mean %.5f max %.5f   |h| rms %.2f max %.1f" % (
                str(dt)[6:], "fp32" if sk is None else str(sk)[6:], e.max(), e.pow(2).mean().sqrt(), tv.mean(), tv.max(), h.pow(2).mean().sqrt(), h.abs().max()))
# the reference's own CUDA arithmetic (torch autocast fp16, yacht/NNet.py:186-193): Linear outputs fp16, SiLU on fp16 tensors
# (fp16 out), LayerNorm computed in fp32 (autocast's fp32 list) -> fp32 out, residual fp32
def fwd_autocast(x):
    dt = torch.float16
    def lin16(x, w, b): return r16(r16(x, dt) @ r16(w, dt).t() + r16(b, dt), dt)
    h = r16(silu(ln(lin16(x, sd["inp.0.weight"], sd["inp.0.bias"]), sd["inp.1.weight"], sd["inp.1.bias"])), torch.float32)
    for i in range(6):
        p = "blocks.%d." % i
        y = ln(r16(silu(lin16(h, sd[p+"fc1.weight"], sd[p+"fc1.bias"])), dt), sd[p+"ln1.weight"], sd[p+"ln1.bias"])
        y = ln(r16(silu(lin16(y, sd[p+"fc2.weight"], sd[p+"fc2.bias"])), dt), sd[p+"ln2.weight"], sd[p+"ln2.bias"])
        h = h + y
    a = r16(silu(ln(h, sd["pi_head.0.weight"], sd["pi_head.0.bias"])), torch.float32)
    return lin16(a, sd["pi_head.2.weight"], sd["pi_head.2.bias"])
with torch.no_grad():
    pi = fwd_autocast(x)
    e = (pi - ref).abs()
    tv = 0.5 * (torch.softmax(ref, -1) - torch.softmax(pi, -1)).abs().sum(-1)
    print("emulated torch autocast fp16 (the reference's CUDA predict): max|dlogit| %.4f  rms %.5f   TV mean %.5f max %.5f" % (e.max(), e.pow(2).mean().sqrt(), tv.mean(), tv.max()))
# not adopted (measured for the split-group experiment): kernel arithmetic v3: fp16 operands, fp16 residual stream, AND the SiLU output / input Linear output rounded to fp16 before LayerNorm
def fwd_v3(x, dt=torch.float16):
    h16 = lambda t: r16(t, torch.float16)
    h = h16(silu(ln(h16(lin(x, sd["inp.0.weight"], sd["inp.0.bias"], dt)), sd["inp.1.weight"], sd["inp.1.bias"])))
    for i in range(6):
        p = "blocks.%d." % i
        y = ln(h16(silu(lin(h, sd[p+"fc1.weight"], sd[p+"fc1.bias"], dt))), sd[p+"ln1.weight"], sd[p+"ln1.bias"])
        y = ln(h16(silu(lin(y, sd[p+"fc2.weight"], sd[p+"fc2.bias"], dt))), sd[p+"ln2.weight"], sd[p+"ln2.bias"])
        h = h16(h + y)
    a = silu(ln(h, sd["pi_head.0.weight"], sd["pi_head.0.bias"]))
    return lin(a, sd["pi_head.2.weight"], sd["pi_head.2.bias"], dt)
with torch.no_grad():
    for dt in (torch.float16, torch.bfloat16):
        pi = fwd_v3(x, dt)
        e = (pi - ref).abs()
        tv = 0.5 * (torch.softmax(ref, -1) - torch.softmax(pi, -1)).abs().sum(-1)
        print("v3 (%s operands; fp16 residual, fp16 SiLU / input-Linear outputs): max|dlogit| %.4f  rms %.5f   TV mean %.5f max %.5f" % (str(dt)[6:], e.max(), e.pow(2).mean().sqrt(), tv.mean(), tv.max()))
