"""Eager definition of the fp8 (e4m3) SageAttention3-style forward.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference names fp8 / FP4 microscaling as the SageAttention3 feature it
does NOT ship (README.md:48-54), so there is no reference code to pin against: parity unpinned.  This file states the
contract the CUDA kernel implements - the int8 pipeline of attention_int8.py:170-257 with e4m3 in place of int8:

  scale = fp16(amax|block| / 448),  value = e4m3(RN(fp16(x / scale)))              (block = 128 rows x D)
  S16   = fp16(float(Q8 K8^T) * sq * sk * qk_scale)        fp16 logits, fp16 running maximum, fp16 subtraction
  P     = exp2(S16 - m'),  sp = exp2(rowmax - m') / 448,   P8 = e4m3(RN(P / sp))    per row per k-tile
  O    += float(P8 V8) * sp * sv                             fp32 accumulation over k-tiles;  O / l, log2-LSE as int8
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504
E4M3_MAX = 448.0


def _e4m3(x: torch.Tensor) -> torch.Tensor:
    """Round to nearest e4m3 (saturating at +-448), returned as fp32 values."""
    return x.float().clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).float()


def quant_block_fp8(x2d: torch.Tensor, blk: int = 128):
    """[N, D] fp16 -> (e4m3 values as fp32 [N, D], raw bytes uint8 [N, D], fp16 scales [N / blk])."""
    n, d = x2d.shape
    assert n % blk == 0
    xb = x2d.reshape(n // blk, blk * d)
    s = torch.amax(xb.abs(), dim=1) / E4M3_MAX                       # fp16
    qv = xb / s[:, None]                                             # fp16 divide
    qv = torch.where(s[:, None] == 0, torch.zeros_like(qv), qv)
    f8 = qv.float().clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return f8.float().reshape(n, d), f8.view(torch.uint8).reshape(n, d), s


def fp8_fwd(q, k, v, blk: int = 128):
    """q, k, v fp16 [B,H,S,D] -> (O fp16 [B,H,S,D], lse32 fp32 [B*H*S], (q8, k8, v8 bytes, sq, sk, sv))."""
    B, H, S, D = q.shape
    N, G = B * H * S, B * H
    qf, qb, sq = quant_block_fp8(q.reshape(N, D), blk)
    kf, kb, sk = quant_block_fp8(k.reshape(N, D), blk)
    vf, vb, sv = quant_block_fp8(v.reshape(N, D), blk)
    qk_scale = (1.0 / math.sqrt(D)) * LOG2E
    qg, kg, vg = qf.view(G, S, D), kf.view(G, S, D), vf.view(G, S, D)
    sq_rows = sq.view(G, S // blk).repeat_interleave(blk, dim=1)[..., None].float()
    sk_g, sv_g = sk.view(G, S // blk), sv.view(G, S // blk)
    O = torch.zeros((G, S, D))
    l = torch.full((G, S, 1), 1.0)
    m = torch.full((G, S, 1), float("-inf"), dtype=torch.float16)
    for j in range(S // blk):
        ks = slice(j * blk, (j + 1) * blk)
        acc = torch.matmul(qg, kg[:, ks].transpose(1, 2))             # exact: e4m3 products summed in fp32 (|sum| small)
        S16 = (acc * (sq_rows * sk_g[:, j].view(G, 1, 1).float() * qk_scale)).to(torch.float16)
        row_max = torch.amax(S16, -1, keepdim=True)
        m_new = torch.max(m, row_max)
        P = torch.exp2((S16 - m_new).float())
        rescale = torch.exp2((m - m_new).float())
        sp = torch.exp2((row_max - m_new).float()) / E4M3_MAX
        l = l * rescale + P.sum(-1, keepdim=True)
        P8 = _e4m3(P / sp)
        O = O * rescale + torch.matmul(P8, vg[:, ks]) * sp * sv_g[:, j].view(G, 1, 1).float()
        m = m_new
    lse32 = m.squeeze(-1).float() + torch.log2(l).squeeze(-1)
    return (O / l).to(torch.float16).view(B, H, S, D), lse32.reshape(N), (qb, kb, vb, sq, sk, sv)
