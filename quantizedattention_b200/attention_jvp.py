"""Drop-in for the reference's `attention_jvp.py` (forward-mode JVP attention), B200-native.

  helion_attention_jvp_forward_fp32(q,k,v,tq,tk,tv) -> (O, tO, lse)     attention_jvp.py:33-195
  baseline_pytorch_attention(q,k,v)                                     attention_jvp.py:197-215
plus (LEDGER J-1, north_star "forward-mode-AD-compatible callables") `jvp_attention(q,k,v)`, a
torch.autograd.Function with a `jvp` staticmethod so that torch.func.jvp / torch.autograd.forward_ad route
through the fused kernel.
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function

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


class _JvpAttention(Function):
    """O = softmax(q k^T / sqrt(d)) v with a fused forward-mode rule (no reverse-mode rule: the reference has none)."""

    @staticmethod
    def forward(q, k, v):
        zeros = torch.zeros_like
        O, _, _ = ops.jvp_fwd(q, k, v, zeros(q), zeros(k), zeros(v))
        return O

    # TODO(perf): a primal-only launch would skip the three tangent contractions; forward() is only reached when the
    # caller asks for O without tangents, the fused O + tO path is jvp() below.

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_forward(*inputs)

    @staticmethod
    def jvp(ctx, tq, tk, tv):
        q, k, v = ctx.saved_tensors
        uq, uk, uv = ops._unwrap(q), ops._unwrap(k), ops._unwrap(v)
        with ops._raw_mode():
            z = lambda t, ref: torch.zeros_like(ref) if t is None else ops._unwrap(t)
            _, tO, _ = ops.jvp_fwd(uq, uk, uv, z(tq, uq), z(tk, uk), z(tv, uv))
        return tO

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("jvp_attention is forward-mode only; use flash_atten_2_bf16 for reverse mode")


def jvp_attention(q, k, v):
    """Forward-mode-AD-compatible attention: `torch.func.jvp(jvp_attention, (q,k,v), (tq,tk,tv))` -> (O, tO)."""
    return _JvpAttention.apply(q, k, v)
