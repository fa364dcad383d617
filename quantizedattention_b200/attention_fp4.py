"""NVFP4 (microscaling) SageAttention3-style forward - the feature the reference's README names as the
SageAttention3 headline but does not implement (README.md:48-54):

  sage_attention_3_fp4(q, k, v, causal=False) -> O fp16 [B,H,S,D]      forward only (inference)
  quantise_fp4(q, k, v)         -> Fp4Operands (codes, scale-factor atoms, per-head scales, k_mean)
  fp4_fwd_prequant(ops)         -> (O fp16 [B,H,S,D], lse fp32 [B*H, S], log2 domain)

Q, K (mean-smoothed) and V are quantised to e2m1 with one e4m3 scale per 16 elements along the contraction axis (D for
Q / K, keys for V, which is stored transposed) and one fp32 scale per head; P is microscaled per row and 16 keys inside
the kernel.  Both contractions run on tcgen05 `kind::mxf4nvf4.block_scale` with fp32 accumulation in TMEM.
D = 128 (64 runs zero-padded to 128 columns); sequence lengths that are not multiples of 128 are zero-padded here and masked in the kernel.  Contract and tolerances: oracle/fp4_ref.py, tests/test_fp4_fwd_gpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib, ops


@dataclass
class Fp4Operands:
    q4: torch.Tensor      # uint8 [B*H*Sq, D/2]
    sfq: torch.Tensor     # uint8 [B*H*Sq/128, D/64, 512]
    sgq: torch.Tensor     # fp32 [B*H]
    k4: torch.Tensor
    sfk: torch.Tensor
    sgk: torch.Tensor
    vt4: torch.Tensor     # uint8 [B*H, D, Sk/2]
    sfv: torch.Tensor     # uint8 [B*H*Sk/128, 2, 512]
    sgv: torch.Tensor
    k_mean: torch.Tensor | None
    shape: tuple          # (B, H, Sq, Sk, D): the PADDED lengths (multiples of 128) the buffers are laid out for
    valid: tuple = None   # (Sq_valid, Sk_valid): the caller's sequence lengths (ragged sequences are zero-padded per head)
    head_dim: int = 128   # the caller's head dimension (64: the operands carry 64 zero columns, sm_scale stays 1/sqrt(64))


def _check(q, k, v):
    for t in (q, k, v):
        if t.dtype != torch.float16:
            raise TypeError("fp4 attention takes fp16 q, k, v")
        if not t.is_cuda:
            raise RuntimeError("fp4 attention needs CUDA tensors (there is no CPU fallback)")
    B, H, Sq, D = q.shape
    Sk = k.shape[2]
    if D not in (64, 128):
        raise ValueError("fp4 attention is built for head dimensions 128 and 64 (64 runs zero-padded to 128)")
    if Sq <= 0 or Sk <= 0:
        raise ValueError("fp4 attention needs non-empty sequences")
    assert k.shape == v.shape and k.shape[:2] == q.shape[:2] and k.shape[3] == D
    return B, H, Sq, Sk, D


def _quant_rows(x, mean, BH, S, D, S_valid=None):
    dev = x.device
    codes = torch.empty((BH * S, D // 2), dtype=torch.uint8, device=dev)
    sf = torch.empty((BH * S // 128, D // 64, 512), dtype=torch.uint8, device=dev)
    sg = torch.empty((BH,), dtype=torch.float32, device=dev)
    ws = torch.empty((2 * BH,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev), ops._timed("fp4_quant_rows"):
        _lib.check(L.qa_fp4_quant_rows_ragged(_lib.ptr(x), _lib.ptr(mean) if mean is not None else None, _lib.ptr(ws), _lib.ptr(codes),
                                              _lib.ptr(sf), _lib.ptr(sg), BH, S, S if S_valid is None else int(S_valid), D,
                                              _lib.cur_stream()), "qa_fp4_quant_rows")
    return codes, sf, sg


def quantise_fp4(q_fp16, k_fp16, v_fp16, smooth_k: bool = True) -> Fp4Operands:
    B, H, Sq_v, Sk_v, D_v = _check(q_fp16, k_fp16, v_fp16)
    BH, D = B * H, 128
    q, k, v = q_fp16.contiguous(), k_fp16.contiguous(), v_fp16.contiguous()
    if D_v != D:                                                      # head dimension 64: zero columns change neither Q K^T nor the
        q, k, v = [torch.nn.functional.pad(t, (0, D - D_v)) for t in (q, k, v)]   # block scales; the extra O columns are dropped
    k_mean = ops.k_mean(k) if smooth_k else None                      # fp16 [B,H,1,D], over the valid keys
    # ragged sequences (the reference's hl.tile clamps its last tile): zero padding per head to a multiple of 128; the padded
    # K rows stay zero after the smoothing and the kernel gives the padded keys weight 0
    Sq, Sk = ops._ceil128(Sq_v), ops._ceil128(Sk_v)
    q, k, v = ops._pad_seq(q, Sq), ops._pad_seq(k, Sk), ops._pad_seq(v, Sk)
    q4, sfq, sgq = _quant_rows(q, None, BH, Sq, D)
    k4, sfk, sgk = _quant_rows(k, k_mean, BH, Sk, D, Sk_v)
    dev = q.device
    vt4 = torch.empty((BH, D, Sk // 2), dtype=torch.uint8, device=dev)
    sfv = torch.empty((BH * Sk // 128, 2, 512), dtype=torch.uint8, device=dev)
    sgv = torch.empty((BH,), dtype=torch.float32, device=dev)
    ws = torch.empty((2 * BH,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev), ops._timed("fp4_quant_vt"):
        _lib.check(L.qa_fp4_quant_vt(_lib.ptr(v), _lib.ptr(ws), _lib.ptr(vt4), _lib.ptr(sfv), _lib.ptr(sgv), BH, Sk, D,
                                     _lib.cur_stream()), "qa_fp4_quant_vt")
    return Fp4Operands(q4, sfq, sgq, k4, sfk, sgk, vt4, sfv, sgv, k_mean, (B, H, Sq, Sk, D), (Sq_v, Sk_v), D_v)


def fp4_fwd_prequant(o: Fp4Operands, variant: int = 0, causal: bool = False):
    """variant 0 (default): one CTA per SM, 128-key tiles, de-phased exp warps; 1: two CTAs per SM, 64-key online-softmax steps."""
    B, H, Sq, Sk, D = o.shape
    Sq_v, Sk_v = o.valid if o.valid is not None else (Sq, Sk)
    if causal and Sq_v != Sk_v:
        raise ValueError("causal fp4 attention is self-attention: Sq == Sk")
    dev = o.q4.device
    O = torch.empty((B * H * Sq, D), dtype=torch.float16, device=dev)
    lse = torch.empty((B * H * Sq,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev), ops._timed("fp4_fwd"):
        _lib.check(L.qa_fp4_fwd_ragged(_lib.ptr(o.q4), _lib.ptr(o.sfq), _lib.ptr(o.sgq), _lib.ptr(o.k4), _lib.ptr(o.sfk), _lib.ptr(o.sgk),
                                       _lib.ptr(o.vt4), _lib.ptr(o.sfv), _lib.ptr(o.sgv), _lib.ptr(O), _lib.ptr(lse), B * H, Sq, Sk, Sk_v, D,
                                       int(variant), 2 if causal else 0, float(o.head_dim) ** -0.5, _lib.cur_stream()), "qa_fp4_fwd")
    O, lse = O.view(B, H, Sq, D), lse.view(B * H, Sq)
    if Sq_v != Sq or o.head_dim != D:
        O, lse = O[:, :, :Sq_v, :o.head_dim].contiguous(), lse[:, :Sq_v].contiguous()
    return O, lse


def sage_attention_3_fp4(q_fp16, k_fp16, v_fp16, causal: bool = False):
    """O = softmax(q k^T / sqrt(d)) v through the NVFP4 pipeline, K smoothed with its per-head token mean.  Forward only:
    the result does not require grad.  causal=True is the STRICT mask of the reference's baseline (key < query; row 0 of a
    head = average over all keys), exactly as attention_int8.sage_attention_3_int8(causal=True)."""
    with torch.no_grad():
        return fp4_fwd_prequant(quantise_fp4(q_fp16, k_fp16, v_fp16), causal=causal)[0]
