"""Test yardstick (not product code): a plain PyTorch forward of a policy/value module as the MCTS leaf evaluator --
float32 (what the CPU reference computes), or a reduced-precision torch forward of the same weights.  The product's
evaluator is the hand-written kernel behind mcts.FusedYachtEvaluator; this class exists so tests can feed the SAME
search kernels with library-computed float32 policies and compare."""
import torch


class TorchEvaluator:
    """One batched torch forward for all leaves; predict semantics of yacht/NNet.py:177-195: pi = softmax(logits) over
    all 3226 actions.

    dtype=torch.bfloat16 runs the whole forward in bf16 from a bf16 copy of the weights (no autocast
    cast kernels per call; LayerNorm still accumulates in fp32 inside its kernel); dtype=None keeps the
    module's own precision (fp32, what the CPU reference computes).
    fused_logits=True (bf16 only) pads the policy head to 3232 outputs (16-byte aligned rows -> the fast
    GEMM path) and hands the raw bf16 logits to ya_mcts_expand_logits, which fuses softmax, masking and
    renormalisation: neither float32 logits nor pi are written to HBM."""
    uniform = False
    PADDED = 3232

    def __init__(self, net, dtype=None, autocast_dtype=None, fused_logits=False):
        import copy
        self.dtype = dtype
        self.autocast_dtype = autocast_dtype
        self.returns_logits = bool(fused_logits)
        self.net = (copy.deepcopy(net).to(dtype) if dtype is not None else net).eval()
        if self.returns_logits:
            assert dtype in (torch.bfloat16, torch.float16), "fused logits need a 16-bit forward"
            head = self.net.pi_head[2]
            padded = torch.nn.Linear(head.in_features, self.PADDED, device=head.weight.device, dtype=head.weight.dtype)
            with torch.no_grad():
                padded.weight.zero_()
                padded.bias.zero_()
                padded.weight[:head.out_features].copy_(head.weight)
                padded.bias[:head.out_features].copy_(head.bias)
            self.net.pi_head[2] = padded

    @torch.no_grad()
    def __call__(self, features, need_eval=None, leaf_states=None):
        if self.dtype is not None:
            logits, v = self.net(features.to(self.dtype))
        elif self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                logits, v = self.net(features)
        else:
            logits, v = self.net(features)
        v = v.float().reshape(-1).contiguous()
        if self.returns_logits:
            return logits.contiguous(), v
        return torch.softmax(logits.float(), dim=1).contiguous(), v
