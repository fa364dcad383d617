"""Layout adapters and `nn.Module` wrappers around the three drop-in paths.

The reference (and the kernels) use contiguous `[B, H, S, D]`; most model code carries `[B, S, H, D]` (what
`F.scaled_dot_product_attention` callers get from a fused QKV projection before their transpose).  The adapters below
accept either layout, run the kernels on a `[B, H, S, D]` contiguous view (one transposing copy per tensor when the input
is `bshd`; none when it already is `bhsd`), and return the output in the caller's layout.  Nothing here computes
attention: every path ends in the C-ABI kernels.
"""
from __future__ import annotations

import torch
from torch import nn

from . import attention_bf16, attention_int8, attention_jvp

_LAYOUTS = ("bhsd", "bshd")


def _to_bhsd(t: torch.Tensor, layout: str) -> torch.Tensor:
    if layout not in _LAYOUTS:
        raise ValueError('layout must be "bhsd" ([B,H,S,D], the reference layout) or "bshd" ([B,S,H,D])')
    if t.dim() != 4:
        raise ValueError("attention inputs are 4-D")
    return t.contiguous() if layout == "bhsd" else t.transpose(1, 2).contiguous()


def _from_bhsd(t: torch.Tensor, layout: str) -> torch.Tensor:
    return t if layout == "bhsd" else t.transpose(1, 2)


def sage_attention_int8(q, k, v, causal: bool = False, layout: str = "bshd", **tunables):
    """`sage_attention_3_int8` (attention_int8.py:434-451) for `[B,S,H,D]` (default) or `[B,H,S,D]` fp16 tensors;
    differentiable.  **tunables: Bq=, Bkv=, rounding= (see attention_int8)."""
    o = attention_int8.sage_attention_3_int8(*[_to_bhsd(t, layout) for t in (q, k, v)], causal=causal, **tunables)
    return _from_bhsd(o, layout)


def flash_attention_bf16(q, k, v, causal: bool = False, layout: str = "bshd"):
    """`flash_atten_2_bf16` (attention_bf16.py:87-105): q, k fp16, v bf16 -> O fp32, in the caller's layout."""
    o = attention_bf16.flash_atten_2_bf16(*[_to_bhsd(t, layout) for t in (q, k, v)], bool(causal))
    return _from_bhsd(o, layout)


def jvp_attention(q, k, v, layout: str = "bshd"):
    """`attention_jvp.jvp_attention` (forward-mode-AD attention, fp32) in the caller's layout."""
    return _from_bhsd(attention_jvp.jvp_attention(*[_to_bhsd(t, layout) for t in (q, k, v)]), layout)


def sage_attention_lowp(q, k, v, precision: str = "fp4", layout: str = "bshd", causal: bool = False):
    """Inference-only SageAttention3-style forward in fp8 (e4m3) or NVFP4 (`precision` = "fp8" | "fp4"; fp4: D = 128) for
    `[B,S,H,D]` (default) or `[B,H,S,D]` fp16 tensors; the result does not require grad.  causal (fp4 only): the strict mask
    of the reference's baseline, see attention_fp4.sage_attention_3_fp4."""
    from . import attention_fp4, attention_fp8
    if precision not in ("fp8", "fp4"):
        raise ValueError('precision must be "fp8" or "fp4"')
    if causal and precision != "fp4":
        raise ValueError("the fp8 forward has no causal mode")
    bhsd = [_to_bhsd(t, layout) for t in (q, k, v)]
    o = attention_fp4.sage_attention_3_fp4(*bhsd, causal=causal) if precision == "fp4" else attention_fp8.sage_attention_3_fp8(*bhsd)
    return _from_bhsd(o, layout)


class SageAttention3LowPrecision(nn.Module):
    """`sage_attention_lowp` as a module (inference): `SageAttention3LowPrecision("fp4")(q, k, v)`."""

    def __init__(self, precision: str = "fp4", layout: str = "bshd", causal: bool = False):
        super().__init__()
        if precision not in ("fp8", "fp4"):
            raise ValueError('precision must be "fp8" or "fp4"')
        self.precision, self.layout, self.causal = precision, layout, causal

    def forward(self, q, k, v):
        return sage_attention_lowp(q, k, v, self.precision, self.layout, self.causal)

    def extra_repr(self):
        return f"precision={self.precision!r}, layout={self.layout!r}, causal={self.causal}"


class SageAttention3Int8(nn.Module):
    """Drop-in attention core: `forward(q, k, v)` -> O (fp16), SageAttention3-style int8 forward and backward."""

    def __init__(self, causal: bool = False, layout: str = "bshd", Bq: int | None = None, Bkv: int | None = None,
                 rounding: str | None = None):
        super().__init__()
        if layout not in _LAYOUTS:
            raise ValueError('layout must be "bhsd" or "bshd"')
        self.causal, self.layout = causal, layout
        self.tunables = {k: v for k, v in (("Bq", Bq), ("Bkv", Bkv), ("rounding", rounding)) if v is not None}

    def forward(self, q, k, v):
        return sage_attention_int8(q, k, v, self.causal, self.layout, **self.tunables)

    def extra_repr(self):
        return f"causal={self.causal}, layout={self.layout!r}, tunables={self.tunables}"


class FlashAttentionBF16(nn.Module):
    """`forward(q_fp16, k_fp16, v_bf16)` -> O fp32: bias-corrected bf16 flash attention with its recompute backward."""

    def __init__(self, causal: bool = False, layout: str = "bshd"):
        super().__init__()
        if layout not in _LAYOUTS:
            raise ValueError('layout must be "bhsd" or "bshd"')
        self.causal, self.layout = causal, layout

    def forward(self, q, k, v):
        return flash_attention_bf16(q, k, v, self.causal, self.layout)

    def extra_repr(self):
        return f"causal={self.causal}, layout={self.layout!r}"


class JvpAttention(nn.Module):
    """`forward(q, k, v)` -> O fp32; under torch.func.jvp / forward_ad the tangent comes from the same fused launch."""

    def __init__(self, layout: str = "bshd"):
        super().__init__()
        if layout not in _LAYOUTS:
            raise ValueError('layout must be "bhsd" or "bshd"')
        self.layout = layout

    def forward(self, q, k, v):
        return jvp_attention(q, k, v, self.layout)


def sage_attention_int8_varlen(q, k, v, cu_seqlens, **tunables):
    """Variable-length (packed) int8 attention, the flash-attn calling convention: q, k, v fp16 `[total_tokens, H, D]`,
    `cu_seqlens` int `[n_seq + 1]` (host tensor / list: cumulative sequence starts), every sequence attends to itself
    (non-causal, like the reference's int8 kernel).  Returns O fp16 `[total_tokens, H, D]`; differentiable.

    Sequences of equal length share one launch; a sequence length need not be a multiple of 128 (ragged last tile,
    attention_int8.py:170,176): `sage_attention_3_int8` pads per head and the kernels give the padding weight 0."""
    cu = [int(x) for x in (cu_seqlens.tolist() if hasattr(cu_seqlens, "tolist") else cu_seqlens)]
    if len(cu) < 2 or cu[0] != 0 or cu[-1] != q.shape[0] or any(b <= a for a, b in zip(cu, cu[1:])):
        raise ValueError("cu_seqlens must start at 0, increase strictly and end at total_tokens")
    if q.dim() != 3 or k.shape != q.shape or v.shape != q.shape:
        raise ValueError("q, k, v must share one [total_tokens, H, D] shape")
    by_len = {}
    for b in range(len(cu) - 1):
        by_len.setdefault(cu[b + 1] - cu[b], []).append(cu[b])
    out = torch.empty_like(q)
    for length, starts in by_len.items():
        idx = torch.cat([torch.arange(s0, s0 + length, device=q.device) for s0 in starts])
        take = lambda t: t.index_select(0, idx).view(len(starts), length, *t.shape[1:]).transpose(1, 2).contiguous()   # [n,H,L,D]
        o = attention_int8.sage_attention_3_int8(take(q), take(k), take(v), **tunables)
        out = out.index_copy(0, idx, o.transpose(1, 2).reshape(len(starts) * length, *q.shape[1:]))
    return out
