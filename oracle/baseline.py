"""fp32 PyTorch attention math of the reference (oracle of intent + CPU baseline of record).

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
from __future__ import annotations

import math

import torch


def baseline_pytorch_attention(q, k, v, head_dim=None, causal=False):
    """Restates attention_int8.py:453-481 / attention_bf16.py:450-478 / attention_jvp.py:197-215.

    softmax(q k^T / sqrt(d)) v in fp32.  `causal` is the reference's STRICT mask: key j is kept
    only if i - j > 0 (diagonal masked); masked logits are filled with the finite value
    -128*ln(2), so row 0 (fully masked) becomes a uniform average over ALL keys.
    """
    if head_dim is None:
        head_dim = q.shape[-1]
    p = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(head_dim)
    if causal:
        sq, sk = p.shape[-2:]
        mask = torch.arange(sq, device=q.device)[:, None] - torch.arange(sk, device=q.device)[None, :]
        fill = -128 * torch.log(torch.tensor([2], device=q.device))
        p = torch.where(mask[None, None] > 0, p, fill)
    p = torch.softmax(p.to(torch.float32), dim=-1).to(torch.float32)
    return torch.matmul(p, v)


def baseline_lse_log2(q, k, causal=False):
    """log2-sum-exp2 of the scaled logits (the quantity the reference kernels store as `lse`,
    attention_bf16.py:288), from the same fp32 math; masked entries excluded except row 0."""
    d = q.shape[-1]
    s = torch.matmul(q.float(), k.float().transpose(2, 3)) * (1.0 / math.sqrt(d)) * 1.44269504
    if causal:
        n = s.shape[-1]
        mask = torch.arange(s.shape[-2])[:, None] - torch.arange(n)[None, :]
        s = torch.where(mask[None, None] > 0, s, torch.tensor(-128.0))
    return torch.logsumexp(s * math.log(2.0), dim=-1) / math.log(2.0)
