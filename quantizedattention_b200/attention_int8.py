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
Bq = Bkv = 128 (the tcgen05 tile); pass Bq= / Bkv= per call (or `set_block_sizes` for the process default): 32/32 is
the reference's untuned default.
"""
from __future__ import annotations

import math
import threading
import warnings

import torch
from torch.autograd import Function

from . import ops

# Process-wide DEFAULTS of the tunables.  Every entry point below also takes them as per-call keyword arguments
# (Bq=, Bkv=, rounding=), which is the re-entrant / thread-safe way to choose them; the setters only change what a call
# without those keywords uses.
_CFG = {"Bq": 128, "Bkv": 128, "nsplit": 0, "rounding": "trunc"}
_CFG_LOCK = threading.Lock()
_BQ_OK, _BKV_OK = (32, 64, 128, 256), (32, 64, 128, 256)


def set_block_sizes(Bq: int = 128, Bkv: int = 128):
    """Default quantisation block sizes (the reference's tunables, attention_int8.py:155-158) for calls that do not pass
    Bq= / Bkv= themselves."""
    if Bq not in _BQ_OK or Bkv not in _BKV_OK:
        raise ValueError("supported tunables: Bq in {32,64,128,256}, Bkv in {32,64,128,256}")
    with _CFG_LOCK:
        _CFG["Bq"], _CFG["Bkv"] = Bq, Bkv


def set_quant_rounding(mode: str = "trunc"):
    """Default int8 rounding of every quantiser of the path (Q, K, V, P, dO, dS).  "trunc" is the reference's
    `.to(torch.int8)` (attention_int8.py:183 etc., LEDGER I-3); "nearest" (round half to even) is the opt-in accuracy
    mode of SURVEY.md 8f.1: it removes the truncation bias, at identical speed, but its int8 tensors are by construction
    not the reference's.  The fused kernels are instantiated for the tuned Bkv = 128 tile in this mode."""
    if mode not in ops.ROUNDING:
        raise ValueError('rounding mode must be "trunc" or "nearest"')
    with _CFG_LOCK:
        _CFG["rounding"] = mode


def _resolve(Bq, Bkv, rounding):
    Bq = _CFG["Bq"] if Bq is None else int(Bq)
    Bkv = _CFG["Bkv"] if Bkv is None else int(Bkv)
    rounding = _CFG["rounding"] if rounding is None else rounding
    if Bq not in _BQ_OK or Bkv not in _BKV_OK:
        raise ValueError("supported tunables: Bq in {32,64,128,256}, Bkv in {32,64,128,256}")
    if rounding not in ops.ROUNDING:
        raise ValueError('rounding mode must be "trunc" or "nearest"')
    return Bq, Bkv, rounding


_CAUSAL_WARNED = False


def _warn_causal():
    global _CAUSAL_WARNED
    if not _CAUSAL_WARNED:
        _CAUSAL_WARNED = True
        warnings.warn("quantizedattention_b200 int8 causal=True follows the reference's own baseline mask "
                      "(attention_int8.py:465-473): STRICT key < query, and row 0 of every head is the uniform average over "
                      "ALL keys (so token 0 depends on future tokens).  This is not the usual key <= query causal mask.",
                      stacklevel=3)


def _check_fp16(*ts):
    for t in ts:
        if t.dtype != torch.float16:
            raise TypeError("int8 attention takes fp16 q, k, v")


def _ceil_to(x: int, m: int) -> int:
    return -(-x // m) * m


def _pad_tokens(t, Sp: int):
    """[B,H,S,D] -> zero-padded [B,H,Sp,D] (a copy only when S != Sp)."""
    B, H, S, D = t.shape
    if S == Sp:
        return t
    out = t.new_zeros((B, H, Sp, D))
    out[:, :, :S] = t
    return out


def _pad_rows(t, BH: int, S: int, Sp: int):
    """Per-head rows: [BH*S, ...] -> zero-padded [BH*Sp, ...]."""
    if S == Sp:
        return t
    rest = t.shape[1:]
    out = t.new_zeros((BH, Sp) + rest)
    out[:, :S] = t.reshape((BH, S) + rest)
    return out.reshape((BH * Sp,) + rest)


def _cut_rows(t, BH: int, S: int, Sp: int):
    """Inverse of _pad_rows: keep the first S of every head's Sp rows."""
    if S == Sp:
        return t
    rest = t.shape[1:]
    return t.reshape((BH, Sp) + rest)[:, :S].reshape((BH * S,) + rest)


def _fwd_tensors(q_fp16, k_fp16, v_fp16, k_mean, Bq, Bkv, rounding, causal):
    """Quantise (K optionally smoothed) and run the fused forward.  Returns everything either wrapper needs.
    Ragged sequence lengths (S not a multiple of 128 / of the block sizes; the reference's hl.tile clamps the last tile,
    attention_int8.py:170,176): the tensors are zero-padded per head, the kernel gives the padded keys weight exactly 0
    (qa_int8_fwd_ragged) and the padding is cut off again: the returned tensors have the caller's S."""
    batch, head, q_tokens, D = q_fp16.shape
    k_tokens = k_fp16.shape[2]
    BH = batch * head
    Sp, Skp = _ceil_to(q_tokens, max(128, Bq)), _ceil_to(k_tokens, max(128, Bkv))
    ragged = (Sp != q_tokens) or (Skp != k_tokens)
    if not ragged:
        q_i8, sq = ops.quant_block(q_fp16, Bq, rounding=rounding)
        k_i8, sk = ops.quant_block(k_fp16, Bkv, mean=k_mean, rows_per_head=k_tokens if k_mean is not None else None, rounding=rounding)
        v_i8, sv = ops.quant_block(v_fp16, Bkv, rounding=rounding)
        O, lse16, lse32 = ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, q_tokens, k_tokens, D, Bq, Bkv,
                                                nsplit=_CFG["nsplit"], want_lse32=True, rounding=rounding, causal=causal)
        return O.view(batch, head, q_tokens, D), lse16, lse32, q_i8, k_i8, v_i8, sq, sk, sv
    if causal:
        raise ValueError("causal int8 attention needs S to be a multiple of 128")
    k_s = k_fp16 if k_mean is None else k_fp16 - k_mean            # fp16, one rounding (attention_int8.py:25); padding stays 0
    q_i8, sq = ops.quant_block(_pad_tokens(q_fp16, Sp), Bq, rounding=rounding)
    k_i8, sk = ops.quant_block(_pad_tokens(k_s, Skp), Bkv, rounding=rounding)
    v_i8, sv = ops.quant_block(_pad_tokens(v_fp16, Skp), Bkv, rounding=rounding)
    O, lse16, lse32 = ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, Sp, Skp, D, Bq, Bkv, nsplit=_CFG["nsplit"],
                                            want_lse32=True, rounding=rounding, sk_valid=k_tokens)
    nbq, nbk = -(-q_tokens // Bq), -(-k_tokens // Bkv)             # blocks per head the reference's clamped tiles produce
    cut = lambda t, S_, Sp_: _cut_rows(t, BH, S_, Sp_).contiguous()
    return (cut(O, q_tokens, Sp).view(batch, head, q_tokens, D), cut(lse16, q_tokens, Sp), cut(lse32, q_tokens, Sp),
            cut(q_i8, q_tokens, Sp), cut(k_i8, k_tokens, Skp), cut(v_i8, k_tokens, Skp),
            cut(sq, nbq, Sp // Bq), cut(sk, nbk, Skp // Bkv), cut(sv, nbk, Skp // Bkv))


def helion_atten_int8_hl_dot_fwd(q_fp16_input, k_fp16_input, v_fp16_input, _want_lse32: bool = False, causal: bool = False,
                                 *, Bq: int | None = None, Bkv: int | None = None, rounding: str | None = None):
    """Quantise Q/K/V per block and run the fused int8 forward.  Returns the reference 10-tuple
    (O fp16 [B,H,S,D], lse fp16 [N], q_i8 [N,D], k_i8_T [D,N], v_i8 [N,D], sq, sk, sv, Bq, Bkv)."""
    batch, head, q_tokens, q_head_dim = q_fp16_input.shape
    _, _, k_tokens, k_head_dim = k_fp16_input.shape
    _, _, v_tokens, v_head_dim = v_fp16_input.shape
    assert k_tokens == v_tokens, "k and v tokens are different"
    assert k_head_dim == v_head_dim, "k head_dim and v head_dim are different"
    _check_fp16(q_fp16_input, k_fp16_input, v_fp16_input)
    Bq, Bkv, rounding = _resolve(Bq, Bkv, rounding)
    if causal:
        _warn_causal()
    O, lse16, lse32, q_i8, k_i8, v_i8, sq, sk, sv = _fwd_tensors(q_fp16_input, k_fp16_input, v_fp16_input, None, Bq, Bkv,
                                                                 rounding, bool(causal))
    out = (O, lse16, q_i8, k_i8.t(), v_i8, sq, sk, sv, Bq, Bkv)
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
                                 v_bh_int8, sv_bh_fp16, O_input_fp16, lse_input_fp16, Bq: int, Bkv: int, causal: bool = False,
                                 *, rounding: str | None = None, kernel: str = "ws", v_fp16=None):
    """Quantised backward (attention_int8.py:268-432 under the 8-LEDGER contract).  Same argument order as the
    reference; `k_mean` is the per-head token mean [B,H,1,D]; `lse` may be the fp16 tensor the forward returned
    or an fp32 copy (LEDGER I-15).  `kernel`: "ws" (warp-specialised, default) or "8warp".
    `v_fp16` (the unquantised V, [B,H,S,D] fp16) selects the SageBwd option: dP = dO V^T is then computed in
    fp16 from the unquantised dO and V instead of the int8 product the reference uses (attention_int8.py:380-384).
    Returns (dq, dk, dv) fp16 [B,H,S,D]."""
    batch, head, q_tokens, head_dim = O_input_fp16.shape
    N = batch * head * q_tokens
    BH, S = batch * head, q_tokens
    assert q_bh_int8.shape == (N, head_dim) and k_bh_int8_T.shape == (head_dim, N), "q/k int8 shapes"
    _, _, rounding = _resolve(Bq, Bkv, rounding)
    k_i8 = k_bh_int8_T.t()
    if not k_i8.is_contiguous():
        k_i8 = k_i8.contiguous()
    dO = dO_input_fp16.contiguous()
    if dO.dtype != torch.float16:
        dO = dO.to(torch.float16)
    O = O_input_fp16.contiguous()
    lse32 = lse_input_fp16.to(torch.float32).contiguous()
    q_i8, v_i8, sq, sk, sv = q_bh_int8.contiguous(), v_bh_int8.contiguous(), sq_bh_fp16, sk_bh_fp16, sv_bh_fp16
    Sp = _ceil_to(S, max(128, Bq, Bkv))
    if Sp != S:                                                    # ragged: back to the zero-padded layout the kernels take
        if causal:
            raise ValueError("causal int8 attention needs S to be a multiple of 128")
        pad = lambda t, S_, Sp_: _pad_rows(t, BH, S_, Sp_)
        dO, O = pad(dO.view(N, head_dim), S, Sp), pad(O.view(N, head_dim), S, Sp)
        q_i8, k_i8, v_i8, lse32 = pad(q_i8, S, Sp), pad(k_i8, S, Sp), pad(v_i8, S, Sp), pad(lse32, S, Sp)
        sq, sk, sv = pad(sq, -(-S // Bq), Sp // Bq), pad(sk, -(-S // Bkv), Sp // Bkv), pad(sv, -(-S // Bkv), Sp // Bkv)
    delta = ops.bwd_delta(dO, O)
    do_i8, s_do = ops.quant_block(dO, Bq, rounding=rounding)
    km = None
    if k_mean_bh_fp16 is not None:
        assert k_mean_bh_fp16.numel() == batch * head * head_dim, "k_mean must be the per-head token mean [B,H,1,D]"
        km = k_mean_bh_fp16.to(torch.float16).contiguous()
    if v_fp16 is not None:                                         # SageBwd: dO V^T stays in fp16
        if causal or Sp != S:
            raise ValueError("the SageBwd option is built for non-causal attention with S a multiple of 128")
        if v_fp16.dtype != torch.float16 or v_fp16.shape != O_input_fp16.shape:
            raise TypeError("v_fp16 must be the fp16 [B,H,S,D] value tensor of the forward")
        dq, dk, dv = ops.int8_bwd_sage(q_i8, k_i8, v_fp16.contiguous().view(N, head_dim), do_i8, dO.view(N, head_dim), sq, sk, s_do,
                                       lse32, delta, km, BH, S, head_dim, Bq, Bkv, rounding=rounding)
        shp = (batch, head, q_tokens, head_dim)
        return dq.view(shp), dk.view(shp), dv.view(shp)
    dq, dk, dv = ops.int8_bwd_prequant(q_i8, k_i8, v_i8, do_i8, sq, sk, sv, s_do, lse32, delta, km, BH, Sp, head_dim,
                                       Bq, Bkv, rounding=rounding, causal=causal, kernel=kernel,
                                       s_valid=S if Sp != S else None)
    shp = (batch, head, q_tokens, head_dim)
    cut = lambda t: _cut_rows(t, BH, S, Sp).contiguous().view(shp)
    return cut(dq), cut(dk), cut(dv)


class _SageInt8Fn(Function):
    """The Function that actually runs: the reference's 11 outputs plus the fp32 log2-LSE as a 12th, non-differentiable
    output, so that the backward recomputes P from an fp32 lse (LEDGER I-15) without any state outside ctx.  Block sizes,
    rounding mode and the mask are explicit arguments."""

    @staticmethod
    def forward(q_fp16, k_fp16, v_fp16, causal, Bq, Bkv, rounding, sage_bwd=False):
        k_mean_fp16 = ops.k_mean(k_fp16)                                       # K-smoothing (LEDGER I-1)
        O, lse16, lse32, q_i8, k_i8, v_i8, sq, sk, sv = _fwd_tensors(q_fp16, k_fp16, v_fp16, k_mean_fp16, Bq, Bkv, rounding, causal)
        return (O, lse16, k_mean_fp16, q_i8, k_i8.t(), v_i8, sq, sk, sv, Bq, Bkv, lse32)

    @staticmethod
    def setup_context(ctx, inputs, output):
        O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, Bq, Bkv, lse32 = output
        ctx.mark_non_differentiable(l_bh_fp16, k_mean_fp16, sq, sk, sv, lse32)
        ctx.set_materialize_grads(False)       # do not allocate zero grads for the auxiliary outputs
        sage_bwd = bool(inputs[7]) if len(inputs) > 7 else False
        ctx.save_for_backward(O_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv,
                              inputs[2] if sage_bwd else None,       # SageBwd keeps the unquantised V for dO V^T
                              lse32)
        ctx.args = (Bq, Bkv, bool(inputs[3]), inputs[6])

    @staticmethod
    def backward(ctx, dO_fp16, *_ignored):
        Bq, Bkv, causal, rounding = ctx.args
        if dO_fp16 is None:
            return (None,) * 8
        O_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, v_sage, lse32 = ctx.saved_tensors
        dq, dk, dv = helion_atten_int8_hl_dot_bwd(dO_fp16, q_bh_int8, sq, k_bh_int8_T, k_mean_fp16, sk, v_bh_int8, sv,
                                                  O_fp16, lse32, Bq, Bkv, causal=causal, rounding=rounding, v_fp16=v_sage)
        return dq, dk, dv, None, None, None, None, None


class SageAttention3_Int8_autograd_function(Function):
    """attention_int8.py:20-95.  `.apply(q, k, v)` -> the reference's 11-tuple
    (O, lse fp16 [N], k_mean [B,H,1,D], q_i8 [N,D], k_i8_T [D,N], v_i8 [N,D], sq, sk, sv, Bq, Bkv), differentiable through
    O.  `.apply` is a thin shim over `_SageInt8Fn` (which carries the fp32 lse to its backward as a hidden 12th output);
    `forward` / `setup_context` / `backward` keep the reference's literal signatures for callers that use them directly
    (that route has only the fp16 lse the reference returns)."""

    @classmethod
    def apply(cls, q_fp16, k_fp16, v_fp16, causal: bool = False, *, Bq: int | None = None, Bkv: int | None = None,
              rounding: str | None = None, sage_bwd: bool = False):
        _check_fp16(q_fp16, k_fp16, v_fp16)
        Bq, Bkv, rounding = _resolve(Bq, Bkv, rounding)
        if causal:
            _warn_causal()
        return _SageInt8Fn.apply(q_fp16, k_fp16, v_fp16, bool(causal), Bq, Bkv, rounding, bool(sage_bwd))[:11]

    @staticmethod
    def forward(q_fp16, k_fp16, v_fp16, causal=False):
        _check_fp16(q_fp16, k_fp16, v_fp16)
        Bq, Bkv, rounding = _resolve(None, None, None)
        return _SageInt8Fn.forward(q_fp16, k_fp16, v_fp16, bool(causal), Bq, Bkv, rounding)[:11]

    @staticmethod
    def setup_context(ctx, inputs, output):
        O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv, Bq, Bkv = output
        ctx.mark_non_differentiable(l_bh_fp16, k_mean_fp16, sq, sk, sv)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv)
        ctx.args = (Bq, Bkv, bool(inputs[3]) if len(inputs) > 3 else False)

    @staticmethod
    def backward(ctx, dO_fp16, *_ignored):
        Bq, Bkv, causal = ctx.args
        pad = (None,) * (len(ctx.needs_input_grad) - 3)
        if dO_fp16 is None:
            return (None, None, None) + pad
        O_fp16, l_bh_fp16, k_mean_fp16, q_bh_int8, k_bh_int8_T, v_bh_int8, sq, sk, sv = ctx.saved_tensors
        dq, dk, dv = helion_atten_int8_hl_dot_bwd(dO_fp16, q_bh_int8, sq, k_bh_int8_T, k_mean_fp16, sk, v_bh_int8, sv,
                                                  O_fp16, l_bh_fp16, Bq, Bkv, causal=causal)
        return (dq, dk, dv) + pad


def sage_attention_3_int8(q_fp16, k_fp16, v_fp16, causal: bool = False, *, Bq: int | None = None, Bkv: int | None = None,
                          rounding: str | None = None, sage_bwd: bool = False):
    """attention_int8.py:434-451: returns O fp16 [B,H,S,D], differentiable w.r.t. q, k, v.
    `causal=True` (absent in the reference's int8 kernel) applies the strict mask of the reference's own
    `baseline_pytorch_attention(..., causal=True)`: key < query, row 0 of a head = uniform average over all keys (a
    one-time warning says so).  Bq / Bkv / rounding: per-call tunables (default: the module defaults).
    sage_bwd=True: the backward keeps dO V^T in fp16 (SageAttention3's SageBwd) instead of quantising it as
    the reference does; it costs the fp16 V in the saved context and runs the (slower) block kernel."""
    if not (torch.is_grad_enabled() and (q_fp16.requires_grad or k_fp16.requires_grad or v_fp16.requires_grad)):
        # inference: the same kernels without the autograd.Function machinery (its per-call argument binding costs more host time
        # than the kernels of a small problem take on the GPU)
        _check_fp16(q_fp16, k_fp16, v_fp16)
        Bq, Bkv, rounding = _resolve(Bq, Bkv, rounding)
        if causal:
            _warn_causal()
        return _SageInt8Fn.forward(q_fp16, k_fp16, v_fp16, bool(causal), Bq, Bkv, rounding)[0]
    return SageAttention3_Int8_autograd_function.apply(q_fp16, k_fp16, v_fp16, causal, Bq=Bq, Bkv=Bkv, rounding=rounding,
                                                       sage_bwd=sage_bwd)[0]
