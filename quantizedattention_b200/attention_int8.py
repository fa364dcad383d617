"""Drop-in for the reference's `attention_int8.py` (SageAttention3-style int8 attention), B200-native.

Same callables, argument order, tensor layouts and output tuples as the reference:
  sage_attention_3_int8(q,k,v)                         attention_int8.py:434-451
  SageAttention3_Int8_autograd_function                attention_int8.py:20-95
  helion_atten_int8_hl_dot_fwd(q, k_smoothed, v)       attention_int8.py:101-262
  helion_atten_int8_hl_dot_bwd(dO, q_i8, sq, k_i8_T, k_mean, sk, v_i8, sv, O, lse, Bq, Bkv)   :268-432
  baseline_pytorch_attention(q,k,v,head_dim,causal)    attention_int8.py:453-481
Every kernel is hand-written sm_100a CUDA behind the C-ABI library (include/qattn.h); there is no
Triton/Helion/CPU path.  Deviations from the reference's literal (partly broken) behaviour follow the
contract in DESIGN.md / SURVEY.md 8-LEDGER: per-(b,h) attention, per-head K token mean, corrected backward.

`Bq` / `Bkv` are the reference's tunables (PowerOfTwoFragment(32, 256, 32), attention_int8.py:155-158): the
quantisation block sizes, returned to the caller and forwarded to backward.  The tuned values here are
Bq = Bkv = 128 (the tcgen05 tile); `set_block_sizes` selects others: the forward runs Bq in {32,64,128,256} and
Bkv in {32,64,128} (32/32 is the reference's untuned default), the backward needs 128/128.
"""
from __future__ import annotations

import math
import weakref

import torch
from torch.autograd import Function

from . import ops

# fp32 copy of the log2-LSE for the backward (LEDGER I-15), keyed by the identity of the fp16 lse tensor the forward
# returns: forward() has no ctx in the new-style Function API, so setup_context() picks the copy up from here.  A
# finalizer drops the entry with its key, so a forward under no_grad leaves nothing behind.
_LSE32 = {}


def _stash_lse32(lse16, lse32):
    k = id(lse16)
    _LSE32[k] = lse32
    weakref.finalize(lse16, _LSE32.pop, k, None)

_CFG = {"Bq": 128, "Bkv": 128, "nsplit": 2, "rounding": "trunc"}


def set_block_sizes(Bq: int = 128, Bkv: int = 128):
    if Bq not in (32, 64, 128, 256) or Bkv not in (32, 64, 128):
        raise ValueError("supported tunables: Bq in {32,64,128,256}, Bkv in {32,64,128} (backward: 128/128)")
    _CFG["Bq"], _CFG["Bkv"] = Bq, Bkv


def set_quant_rounding(mode: str = "trunc"):
    """int8 rounding of every quantiser of the path (Q, K, V, P, dO, dS).  "trunc" is the reference's `.to(torch.int8)`
    (attention_int8.py:183 etc., LEDGER I-3) and the default; "nearest" (round half to even) is the opt-in accuracy mode
    of SURVEY.md 8f.1: it removes the truncation bias, at identical speed, but its int8 tensors are by construction
    not the reference's.  The fused kernels are instantiated for the tuned Bkv = 128 tile in this mode."""
    if mode not in ops.ROUNDING:
        raise ValueError('rounding mode must be "trunc" or "nearest"')
    _CFG["rounding"] = mode


def helion_atten_int8_hl_dot_fwd(q_fp16_input, k_fp16_input, v_fp16_input, _want_lse32: bool = False, causal: bool = False):
    """Quantise Q/K/V per block and run the fused int8 forward.  Returns the reference 10-tuple
    (O fp16 [B,H,S,D], lse fp16 [N], q_i8 [N,D], k_i8_T [D,N], v_i8 [N,D], sq, sk, sv, Bq, Bkv)."""
    batch, head, q_tokens, q_head_dim = q_fp16_input.shape
    _, _, k_tokens, k_head_dim = k_fp16_input.shape
    _, _, v_tokens, v_head_dim = v_fp16_input.shape
    assert k_tokens == v_tokens, "k and v tokens are different"
    assert k_head_dim == v_head_dim, "k head_dim and v head_dim are different"
    for t in (q_fp16_input, k_fp16_input, v_fp16_input):
        if t.dtype != torch.float16:
            raise TypeError("int8 attention takes fp16 q, k, v")
    Bq, Bkv = _CFG["Bq"], _CFG["Bkv"]
    D = q_head_dim
    rnd = _CFG["rounding"]
    q_i8, sq = ops.quant_block(q_fp16_input, Bq, rounding=rnd)
    k_i8, sk = ops.quant_block(k_fp16_input, Bkv, rounding=rnd)
    v_i8, sv = ops.quant_block(v_fp16_input, Bkv, rounding=rnd)
    O, lse16, lse32 = ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, batch * head, q_tokens, k_tokens, D, Bq, Bkv,
                                            nsplit=_CFG["nsplit"], want_lse32=_want_lse32, rounding=rnd, causal=causal)
    out = (O.view(batch, head, q_tokens, D), lse16, q_i8, k_i8.t(), v_i8, sq, sk, sv, Bq, Bkv)
    return out + (lse32,) if _want_lse32 else out


def baseline_pytorch_attention(q, k, v, head_dim, causal):
    """The reference's fp32 PyTorch attention math (strict causal mask, finite -128*ln2 fill)."""
    p = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(head_dim)
    if causal:
        b, h, q_token, k_token = p.shape
        mask = torch.arange(q_token, device=q.device)[:, None] - torch.arange(k_token, device=q.device)[None, :]
        p = torch.where(mask[None, None] > 0, p, -128 * torch.log(torch.tensor([2], device=q.device)))
    p = torch.softmax(p.to(torch.float32), dim=-1).to(torch.float32)
    return torch.matmul(p, v)


def helion_atten_int8_hl_dot_bwd(dO_input_fp16, q_bh_int8, sq_bh_fp16, k_bh_int8_T, k_mean_bh_fp16, sk_bh_fp16,
                                 v_bh_int8, sv_bh_fp16, O_input_fp16, lse_input_fp16, Bq: int, Bkv: int, causal: bool = False):
    """Quantised backward (attention_int8.py:268-432 under the 8-LEDGER contract).  Same argument order as the
    reference; `k_mean` is the per-head token mean [B,H,1,D]; `lse` may be the fp16 tensor the forward returned
    or an fp32 copy (LEDGER I-15).  Returns (dq, dk, dv) fp16 [B,H,S,D]."""
    batch, head, q_tokens, head_dim = O_input_fp16.shape
    N = batch * head * q_tokens
    assert q_bh_int8.shape == (N, head_dim) and k_bh_int8_T.shape == (head_dim, N), "q/k int8 shapes"
    if Bq != 128 or Bkv != 128:
        raise ValueError("backward needs Bq = Bkv = 128 (int32 accumulation depth, SURVEY.md 7 hard part 3)")
    k_i8 = k_bh_int8_T.t()
    if not k_i8.is_contiguous():
        k_i8 = k_i8.contiguous()
    dO = dO_input_fp16.contiguous()
    if dO.dtype != torch.float16:
        dO = dO.to(torch.float16)
    delta = ops.bwd_delta(dO, O_input_fp16)
    do_i8, s_do = ops.quant_block(dO, Bq, rounding=_CFG["rounding"])
    lse32 = lse_input_fp16.to(torch.float32).contiguous()
    km = None
    if k_mean_bh_fp16 is not None:
        assert k_mean_bh_fp16.numel() == batch * head * head_dim, "k_mean must be the per-head token mean [B,H,1,D]"
        km = k_mean_bh_fp16.to(torch.float16).contiguous()
    dq, dk, dv = ops.int8_bwd_prequant(q_bh_int8.contiguous(), k_i8, v_bh_int8.contiguous(), do_i8, sq_bh_fp16,
                                       sk_bh_fp16, sv_bh_fp16, s_do, lse32, delta, km, batch * head, q_tokens, head_dim,
                                       Bq, Bkv, rounding=_CFG["rounding"], causal=causal)
    shp = (batch, head, q_tokens, head_dim)
    return dq.view(shp), dk.view(shp), dv.view(shp)


class SageAttention3_Int8_autograd_function(Function):
    """attention_int8.py:20-95.  forward(q,k,v) -> 11-tuple
    (O, lse fp16 [N], k_mean [B,H,1,D], q_i8 [N,D], k_i8_T [D,N], v_i8 [N,D], sq, sk, sv, Bq, Bkv)."""

    @staticmethod
    def forward(q_fp16, k_fp16, v_fp16, causal=False):
        for t in (q_fp16, k_fp16, v_fp16):
            if t.dtype != torch.float16:
                raise TypeError("int8 attention takes fp16 q, k, v")
        batch, head, q_tokens, D = q_fp16.shape
        k_tokens = k_fp16.shape[2]
        Bq, Bkv = _CFG["Bq"], _CFG["Bkv"]
        k_mean_fp16 = ops.k_mean(k_fp16)                                       # K-smoothing (LEDGER I-1)
        rnd = _CFG["rounding"]
        q_i8, sq = ops.quant_block(q_fp16, Bq, rounding=rnd)
        k_i8, sk = ops.quant_block(k_fp16, Bkv, mean=k_mean_fp16, rows_per_head=k_tokens, rounding=rnd)   # fused k - mean
        v_i8, sv = ops.quant_block(v_fp16, Bkv, rounding=rnd)
        O, lse16, lse32 = ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, batch * head, q_tokens, k_tokens, D, Bq, Bkv,
                                                nsplit=_CFG["nsplit"], want_lse32=True, rounding=rnd, causal=bool(causal))
        _stash_lse32(lse16, lse32)                                             # picked up by setup_context
        return (O.view(batch, head, q_tokens, D), lse16, k_mean_fp16, q_i8, k_i8.t(), v_i8, sq, sk, sv, Bq, Bkv)

    @staticmethod
    def setup_context(ctx, inputs, output):
        O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, Bq, Bkv = output
        ctx.mark_non_differentiable(l_bh_fp16, k_mean_fp16, sq, sk, sv)
        ctx.set_materialize_grads(False)       # do not allocate zero grads for the 10 auxiliary outputs
        lse32 = _LSE32.pop(id(l_bh_fp16), None)
        ctx.save_for_backward(O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, lse32)
        ctx.args = (Bq, Bkv, bool(inputs[3]) if len(inputs) > 3 else False)

    @staticmethod
    def backward(ctx, dO_fp16, *_ignored):
        Bq, Bkv, causal = ctx.args
        pad = (None,) * (len(ctx.needs_input_grad) - 3)
        if dO_fp16 is None:
            return (None, None, None) + pad
        O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, lse32 = ctx.saved_tensors
        dq, dk, dv = helion_atten_int8_hl_dot_bwd(dO_fp16, q_bh_int8, sq, k_bh_int8_T, k_mean_fp16, sk, v_bh_int8, sv,
                                                  O_fp16, lse32 if lse32 is not None else l_bh_fp16, Bq, Bkv, causal=causal)
        return (dq, dk, dv) + pad


def sage_attention_3_int8(q_fp16, k_fp16, v_fp16, causal: bool = False):
    """attention_int8.py:434-451: returns O fp16 [B,H,S,D], differentiable w.r.t. q, k, v.
    `causal=True` (absent in the reference's int8 kernel) applies the strict mask of the reference's own
    `baseline_pytorch_attention(..., causal=True)`: key < query, row 0 of a head = uniform average over all keys."""
    if causal:
        return SageAttention3_Int8_autograd_function.apply(q_fp16, k_fp16, v_fp16, True)[0]
    return SageAttention3_Int8_autograd_function.apply(q_fp16, k_fp16, v_fp16)[0]
