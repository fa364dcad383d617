"""Drop-in for the reference's `attention_jvp.py` (forward-mode JVP attention), B200-native.

  helion_attention_jvp_forward_fp32(q,k,v,tq,tk,tv) -> (O, tO, lse)     attention_jvp.py:33-195
  baseline_pytorch_attention(q,k,v)                                     attention_jvp.py:197-215
plus (LEDGER J-1, north_star "forward-mode-AD-compatible callables") `jvp_attention(q,k,v)`: under torch.func.jvp /
torch.autograd.forward_ad it takes the tangents off its dual inputs and returns a dual output from ONE fused launch.
"""
from __future__ import annotations

import math

import torch

from . import ops


def helion_attention_jvp_forward_fp32(q_fp32_input, k_fp32_input, v_fp32_input, tan_q_fp32_input, tan_k_fp32_input,
                                      tan_v_fp32_input):
    batch, head, q_tokens, q_head_dim = q_fp32_input.shape
    _, _, k_tokens, k_head_dim = k_fp32_input.shape
    _, _, v_tokens, v_head_dim = v_fp32_input.shape
    assert k_tokens == v_tokens, "input k_tokens must match v_tokens"
    assert q_head_dim == k_head_dim == v_head_dim, "all head dimensions must match for q, k, v tensors"
    return ops.jvp_fwd(q_fp32_input, k_fp32_input, v_fp32_input, tan_q_fp32_input, tan_k_fp32_input, tan_v_fp32_input)


def baseline_pytorch_attention(q, k, v):
    batch, head, tokens, head_dim = q.shape
    s = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(head_dim)
    p = torch.softmax(s.to(torch.float32), dim=-1).to(torch.float32)
    return torch.matmul(p, v)


def jvp_attention(q, k, v):
    """Forward-mode-AD-compatible attention O = softmax(q k^T / sqrt(d)) v (LEDGER J-1).

        O, tO = torch.func.jvp(jvp_attention, (q, k, v), (tq, tk, tv))
        with fwAD.dual_level(): out = jvp_attention(fwAD.make_dual(q, tq), ...);  O, tO = fwAD.unpack_dual(out)

    ONE launch of the fused kernel produces both O and tO (attention_jvp.py:129-190 carries the tangents through the same
    tiles): the inputs' tangents are taken off the dual tensors here, the kernel runs below any functorch level, and the
    result goes back as a dual tensor.  Inputs without a tangent count as zero tangents; without any tangent the call is
    a plain forward.  There is no reverse-mode rule (the reference has none): use flash_atten_2_bf16 for training."""
    import torch.autograd.forward_ad as fwAD
    duals = [fwAD.unpack_dual(t) for t in (q, k, v)]
    tangents = [d.tangent for d in duals]
    if all(t is None for t in tangents):
        zeros = torch.zeros_like
        O, _, _ = ops.jvp_fwd(q, k, v, zeros(q), zeros(k), zeros(v))
        return O
    primals = [d.primal for d in duals]
    tangents = [torch.zeros_like(p) if t is None else t for p, t in zip(primals, tangents)]
    O, tO, _ = ops.jvp_fwd(*[p.detach() for p in primals], *[t.detach() for t in tangents])   # unwraps functorch levels itself
    return fwAD.make_dual(O, tO)
