"""fp8 (e4m3) SageAttention3-style forward - the feature the reference's README names as the
SageAttention3 headline but does not implement (README.md:48-54), built on the same pipeline as the int8 forward:

  sage_attention_3_fp8(q, k, v) -> O fp16 [B,H,S,D]                    forward only (inference)
  helion_atten_fp8_fwd(q, k_smoothed, v) -> (O, lse fp16 [N], q8, k8_T, v8, sq, sk, sv, 128, 128)

Q, K (mean-smoothed) and V are quantised per 128-row block to e4m3 (scale = amax / 448), P per row per k-tile; both
contractions run on tcgen05 `kind::f8f6f4` with fp32 accumulation in TMEM.  Same layouts as attention_int8; the quantised
tensors are returned as torch.float8_e4m3fn views.  S must be a multiple of 128, D in {64, 128}.
"""
from __future__ import annotations

import torch

from . import _lib, ops


def _quant(x, mean=None, rows_per_head=None):
    q, s = ops.quant_block(x, 128, mean=mean, rows_per_head=rows_per_head, rounding="e4m3")
    return q, s


def _fwd(q_fp16, k_fp16, v_fp16, k_mean):
    for t in (q_fp16, k_fp16, v_fp16):
        if t.dtype != torch.float16:
            raise TypeError("fp8 attention takes fp16 q, k, v")
    B, H, Sq, D = q_fp16.shape
    Sk = k_fp16.shape[2]
    if Sq % 128 or Sk % 128:
        raise ValueError("fp8 attention needs sequence lengths that are multiples of 128")
    q8, sq = _quant(q_fp16)
    k8, sk = _quant(k_fp16, k_mean, Sk if k_mean is not None else None)
    v8, sv = _quant(v_fp16)
    dev = q_fp16.device
    O = torch.empty((B * H * Sq, D), dtype=torch.float16, device=dev)
    lse16 = torch.empty((B * H * Sq,), dtype=torch.float16, device=dev)
    lse32 = torch.empty((B * H * Sq,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev), ops._timed("fp8_fwd"):
        _lib.check(L.qa_fp8_fwd(_lib.ptr(q8), _lib.ptr(k8), _lib.ptr(v8), _lib.ptr(sq), _lib.ptr(sk), _lib.ptr(sv), _lib.ptr(O),
                                _lib.ptr(lse16), _lib.ptr(lse32), B * H, Sq, Sk, D, _lib.cur_stream()), "qa_fp8_fwd")
    f8 = lambda t: t.view(torch.float8_e4m3fn)
    return O.view(B, H, Sq, D), lse16, lse32, f8(q8), f8(k8), f8(v8), sq, sk, sv


def helion_atten_fp8_fwd(q_fp16_input, k_fp16_input, v_fp16_input):
    """Same tuple layout as attention_int8.helion_atten_int8_hl_dot_fwd, with e4m3 tensors; K is taken as given."""
    O, lse16, _lse32, q8, k8, v8, sq, sk, sv = _fwd(q_fp16_input, k_fp16_input, v_fp16_input, None)
    return O, lse16, q8, k8.t(), v8, sq, sk, sv, 128, 128


def sage_attention_3_fp8(q_fp16, k_fp16, v_fp16):
    """O = softmax(q k^T / sqrt(d)) v through the fp8 pipeline, K smoothed with its per-head token mean.  Forward only:
    the result does not require grad."""
    for t in (q_fp16, k_fp16, v_fp16):
        if t.dtype != torch.float16:
            raise TypeError("fp8 attention takes fp16 q, k, v")
    if q_fp16.shape[2] % 128 or k_fp16.shape[2] % 128:
        raise ValueError("fp8 attention needs sequence lengths that are multiples of 128")
    with torch.no_grad():
        k_mean = ops.k_mean(k_fp16)
        return _fwd(q_fp16, k_fp16, v_fp16, k_mean)[0]
