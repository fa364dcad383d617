"""Drop-in for the reference's `attention_bf16.py` (bf16 flash attention with the Qiu & Yao softmax bias
correction, fp32-accumulated recompute backward), B200-native.

  flash_atten_2_bf16(q_fp16, k_fp16, v_bf16, causal) -> O fp32                     attention_bf16.py:87-105
  FlashAttention_2_BF16_autograd_function.apply(q,k,v,causal) -> (O, lse)           attention_bf16.py:16-85
  helion_atten_bf16_fwd_training(q,k,v,causal) -> (O fp32, lse fp32 [B*H,S])       attention_bf16.py:111-296
  helion_flash_atten_2_algo_4_bwd(q,k,v,O,lse,causal,dO) -> (dq,dk,dv) fp32         attention_bf16.py:309-448
  baseline_pytorch_attention(q,k,v,head_dim,causal)                                 attention_bf16.py:450-478
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops
from .attention_int8 import baseline_pytorch_attention  # noqa: F401  (identical in both reference files)


def _check_dtypes(q, k, v):
    # LEDGER B-11: the reference fixes dtypes by naming convention only; we accept exactly (fp16, fp16, bf16)
    if q.dtype != torch.float16 or k.dtype != torch.float16 or v.dtype != torch.bfloat16:
        raise TypeError("bf16 attention takes q, k in fp16 and v in bf16")


def helion_atten_bf16_fwd_training(q_fp16_input, k_fp16_input, v_bf16_input, causal: bool):
    batch, head, q_tokens, q_head_dim = q_fp16_input.shape
    k_tokens, v_tokens = k_fp16_input.shape[2], v_bf16_input.shape[2]
    assert k_tokens == v_tokens, "input k_tokens must match v_tokens"
    assert q_head_dim == k_fp16_input.size(-1) == v_bf16_input.size(-1), "all head dimensions must match for q, k, v tensors"
    _check_dtypes(q_fp16_input, k_fp16_input, v_bf16_input)
    return ops.bf16_fwd(q_fp16_input, k_fp16_input, v_bf16_input, bool(causal))


def helion_flash_atten_2_algo_4_bwd(q_input, k_input, v_input, O_input, lse_input, causal: bool, dO_input):
    _check_dtypes(q_input, k_input, v_input)
    return ops.bf16_bwd(q_input, k_input, v_input, O_input, lse_input, bool(causal), dO_input)


class FlashAttention_2_BF16_autograd_function(Function):
    @staticmethod
    def forward(q_fp16, k_fp16, v_bf16, causal):
        return helion_atten_bf16_fwd_training(q_fp16, k_fp16, v_bf16, causal)

    @staticmethod
    def setup_context(ctx, inputs, output):
        q_fp16, k_fp16, v_bf16, causal = inputs
        O_fp32, lse_fp32 = output
        ctx.mark_non_differentiable(lse_fp32)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(q_fp16, k_fp16, v_bf16, O_fp32, lse_fp32)
        ctx.args = causal

    @staticmethod
    def backward(ctx, dO, _lse):
        if dO is None:
            return None, None, None, None
        q_fp16, k_fp16, v_bf16, O_fp32, lse_fp32 = ctx.saved_tensors
        dq, dk, dv = helion_flash_atten_2_algo_4_bwd(q_fp16, k_fp16, v_bf16, O_fp32, lse_fp32, ctx.args, dO)
        return dq, dk, dv, None      # fp32 grads; autograd casts them to the input dtypes (LEDGER B-10)


def flash_atten_2_bf16(q_fp16, k_fp16, v_bf16, causal):
    o_fp32, _lse_fp32 = FlashAttention_2_BF16_autograd_function.apply(q_fp16, k_fp16, v_bf16, causal)
    return o_fp32
