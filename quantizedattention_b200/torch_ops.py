"""`torch.library` registration of the fused paths.

The drop-in modules call the C ABI through ctypes inside `torch.autograd.Function`s, which Dynamo cannot trace.  The
operators below wrap the same calls as opaque `torch.library.custom_op`s with fake-tensor (shape / dtype) kernels and
registered backward formulas, so the attention calls survive `torch.compile(fullgraph=True)` / `torch.export`:

    qattn::sage_int8_fwd / qattn::sage_int8_bwd   ->  sage_attention_3_int8_op(q, k, v)
    qattn::flash_bf16_fwd / qattn::flash_bf16_bwd ->  flash_atten_2_bf16_op(q, k, v, causal)
    qattn::jvp_fwd                                 ->  attention_jvp_op(q, k, v, tq, tk, tv) -> (O, tO, lse)
    qattn::sage_fp8_fwd / qattn::sage_fp4_fwd      ->  sage_attention_3_fp8_op / sage_attention_3_fp4_op(q, k, v)   (inference)

Numerics are those of `attention_int8.sage_attention_3_int8` / `attention_bf16.flash_atten_2_bf16` (same kernels, the
block sizes and rounding mode current at call time are baked in as integer arguments).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import attention_bf16 as _bf16
from . import attention_int8 as _int8
from . import ops

_T = torch.Tensor


# ------------------------------------------------------------------------------------------------ int8 (SageAttention3)
@torch.library.custom_op("qattn::sage_int8_fwd", mutates_args=())
def sage_int8_fwd(q: _T, k: _T, v: _T, Bq: int, Bkv: int, nearest: bool) -> Tuple[_T, _T, _T, _T, _T, _T, _T, _T, _T]:
    """-> (O fp16 [B,H,S,D], lse fp32 [N], k_mean fp16 [B,H,1,D], q_i8, k_i8, v_i8 [N,D], sq, sk, sv fp16)."""
    B, H, S, D = q.shape
    rnd = "nearest" if nearest else "trunc"
    km = ops.k_mean(k)
    q_i8, sq = ops.quant_block(q, Bq, rounding=rnd)
    k_i8, sk = ops.quant_block(k, Bkv, mean=km, rows_per_head=k.shape[2], rounding=rnd)
    v_i8, sv = ops.quant_block(v, Bkv, rounding=rnd)
    O, _lse16, lse32 = ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, B * H, S, k.shape[2], D, Bq, Bkv,
                                             want_lse32=True, rounding=rnd)
    return O.view(B, H, S, D), lse32, km, q_i8, k_i8, v_i8, sq, sk, sv


@sage_int8_fwd.register_fake
def _(q, k, v, Bq, Bkv, nearest):
    B, H, S, D = q.shape
    N, Nk = B * H * S, B * H * k.shape[2]
    i8 = lambda n: q.new_empty((n, D), dtype=torch.int8)
    sc = lambda n: q.new_empty((n,), dtype=torch.float16)
    return (q.new_empty((B, H, S, D), dtype=torch.float16), q.new_empty((N,), dtype=torch.float32),
            q.new_empty((B, H, 1, D), dtype=torch.float16), i8(N), i8(Nk), i8(Nk), sc(N // Bq), sc(Nk // Bkv), sc(Nk // Bkv))


@torch.library.custom_op("qattn::sage_int8_bwd", mutates_args=())
def sage_int8_bwd(dO: _T, O: _T, lse: _T, k_mean: _T, q_i8: _T, k_i8: _T, v_i8: _T, sq: _T, sk: _T, sv: _T, Bq: int,
                  Bkv: int, nearest: bool) -> Tuple[_T, _T, _T]:
    B, H, S, D = O.shape
    rnd = "nearest" if nearest else "trunc"
    dO = dO.to(torch.float16).contiguous()
    delta = ops.bwd_delta(dO, O)
    do_i8, s_do = ops.quant_block(dO, Bq, rounding=rnd)
    dq, dk, dv = ops.int8_bwd_prequant(q_i8, k_i8, v_i8, do_i8, sq, sk, sv, s_do, lse, delta, k_mean, B * H, S, D, Bq, Bkv,
                                       rounding=rnd)
    return dq.view(B, H, S, D), dk.view(B, H, S, D), dv.view(B, H, S, D)


@sage_int8_bwd.register_fake
def _(dO, O, lse, k_mean, q_i8, k_i8, v_i8, sq, sk, sv, Bq, Bkv, nearest):
    return torch.empty_like(O), torch.empty_like(O), torch.empty_like(O)


def _sage_setup(ctx, inputs, output):
    _q, _k, _v, Bq, Bkv, nearest = inputs
    ctx.save_for_backward(*output)
    ctx.cfg = (Bq, Bkv, nearest)


def _sage_backward(ctx, dO, *_unused):
    O, lse, km, q_i8, k_i8, v_i8, sq, sk, sv = ctx.saved_tensors
    dq, dk, dv = sage_int8_bwd(dO, O, lse, km, q_i8, k_i8, v_i8, sq, sk, sv, *ctx.cfg)
    return dq, dk, dv, None, None, None


sage_int8_fwd.register_autograd(_sage_backward, setup_context=_sage_setup)


def sage_attention_3_int8_op(q_fp16: _T, k_fp16: _T, v_fp16: _T, *, Bq: int | None = None, Bkv: int | None = None,
                             rounding: str | None = None) -> _T:
    """`sage_attention_3_int8` (attention_int8.py:434-451) as a traceable operator: O fp16 [B,H,S,D].  Bq / Bkv / rounding
    are per-call arguments (default: the module defaults of attention_int8 at call / trace time)."""
    Bq, Bkv, rounding = _int8._resolve(Bq, Bkv, rounding)
    return sage_int8_fwd(q_fp16, k_fp16, v_fp16, Bq, Bkv, rounding == "nearest")[0]


# ------------------------------------------------------------------------------------------------ bf16 flash attention
@torch.library.custom_op("qattn::flash_bf16_fwd", mutates_args=())
def flash_bf16_fwd(q: _T, k: _T, v: _T, causal: bool) -> Tuple[_T, _T]:
    return ops.bf16_fwd(q, k, v, causal)


@flash_bf16_fwd.register_fake
def _(q, k, v, causal):
    B, H, S, D = q.shape
    return q.new_empty((B, H, S, D), dtype=torch.float32), q.new_empty((B * H, S), dtype=torch.float32)


@torch.library.custom_op("qattn::flash_bf16_bwd", mutates_args=())
def flash_bf16_bwd(q: _T, k: _T, v: _T, O: _T, lse: _T, causal: bool, dO: _T) -> Tuple[_T, _T, _T]:
    return _bf16.helion_flash_atten_2_algo_4_bwd(q, k, v, O, lse, causal, dO)


@flash_bf16_bwd.register_fake
def _(q, k, v, O, lse, causal, dO):
    f32 = lambda t: t.new_empty(t.shape, dtype=torch.float32)
    return f32(q), f32(k), f32(v)


def _flash_setup(ctx, inputs, output):
    q, k, v, causal = inputs
    O, lse = output
    ctx.save_for_backward(q, k, v, O, lse)
    ctx.causal = causal


def _flash_backward(ctx, dO, _dlse):
    q, k, v, O, lse = ctx.saved_tensors
    dq, dk, dv = flash_bf16_bwd(q, k, v, O, lse, ctx.causal, dO.to(torch.float32).contiguous())
    return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None


flash_bf16_fwd.register_autograd(_flash_backward, setup_context=_flash_setup)


def flash_atten_2_bf16_op(q_fp16: _T, k_fp16: _T, v_bf16: _T, causal: bool) -> _T:
    """`flash_atten_2_bf16` (attention_bf16.py:87-105) as a traceable operator: O fp32 [B,H,S,D]."""
    return flash_bf16_fwd(q_fp16, k_fp16, v_bf16, causal)[0]


# ------------------------------------------------------------------------------------------------ forward-mode JVP attention
@torch.library.custom_op("qattn::jvp_fwd", mutates_args=())
def jvp_fwd(q: _T, k: _T, v: _T, tq: _T, tk: _T, tv: _T) -> Tuple[_T, _T, _T]:
    """helion_attention_jvp_forward_fp32 (attention_jvp.py:33-195): (O, tO fp32 [B,H,S,D], lse fp32 [B*H,S])."""
    return ops.jvp_fwd(q, k, v, tq, tk, tv)


@jvp_fwd.register_fake
def _(q, k, v, tq, tk, tv):
    B, H, S, D = q.shape
    f32 = lambda *sh: q.new_empty(sh, dtype=torch.float32)
    return f32(B, H, S, D), f32(B, H, S, D), f32(B * H, S)


def attention_jvp_op(q: _T, k: _T, v: _T, tq: _T, tk: _T, tv: _T) -> Tuple[_T, _T, _T]:
    """The reference's JVP kernel function as a traceable operator (survives torch.compile(fullgraph=True))."""
    return jvp_fwd(q, k, v, tq, tk, tv)


# ------------------------------------------------------------------------------------------------ fp8 / NVFP4 forwards (inference)
@torch.library.custom_op("qattn::sage_fp8_fwd", mutates_args=())
def sage_fp8_fwd(q: _T, k: _T, v: _T) -> _T:
    """attention_fp8.sage_attention_3_fp8: O fp16 [B,H,S,D] through the e4m3 pipeline (no gradient)."""
    from . import attention_fp8
    return attention_fp8.sage_attention_3_fp8(q, k, v)


@sage_fp8_fwd.register_fake
def _(q, k, v):
    return q.new_empty(q.shape, dtype=torch.float16)


@torch.library.custom_op("qattn::sage_fp4_fwd", mutates_args=())
def sage_fp4_fwd(q: _T, k: _T, v: _T, causal: bool) -> _T:
    """attention_fp4.sage_attention_3_fp4: O fp16 [B,H,S,D] through the NVFP4 pipeline (D = 128; no gradient)."""
    from . import attention_fp4
    return attention_fp4.sage_attention_3_fp4(q, k, v, causal)


@sage_fp4_fwd.register_fake
def _(q, k, v, causal):
    return q.new_empty(q.shape, dtype=torch.float16)


def sage_attention_3_fp8_op(q_fp16: _T, k_fp16: _T, v_fp16: _T) -> _T:
    return sage_fp8_fwd(q_fp16, k_fp16, v_fp16)


def sage_attention_3_fp4_op(q_fp16: _T, k_fp16: _T, v_fp16: _T, causal: bool = False) -> _T:
    return sage_fp4_fwd(q_fp16, k_fp16, v_fp16, causal)
