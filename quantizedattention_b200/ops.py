"""Low-level Python bindings of the C-ABI kernels (one function per `qa_*` entry point).

Torch owns every buffer (inputs, outputs, workspaces are `tensor.data_ptr()`); the library never
allocates device memory and never synchronises.  Everything runs on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("quantizedattention_b200 has no CPU path: tensors must live on a CUDA device")


def k_mean(k: torch.Tensor) -> torch.Tensor:
    """Per-(b,h) token mean of K, fp32 accumulate -> fp16 [B,H,1,D] (K-smoothing, LEDGER I-1)."""
    _need_cuda(k)
    assert k.dtype == torch.float16 and k.dim() == 4
    k = k.contiguous()
    B, H, S, D = k.shape
    L = _lib.lib()
    ws = torch.empty(L.qa_k_mean_workspace_bytes(B, H, S, D), dtype=torch.uint8, device=k.device)
    out = torch.empty((B, H, 1, D), dtype=torch.float16, device=k.device)
    with torch.cuda.device(k.device):
        _lib.check(L.qa_k_mean(_lib.ptr(k), _lib.ptr(out), _lib.ptr(ws), ws.numel(), B, H, S, D, _lib.cur_stream()), "qa_k_mean")
    return out


def quant_block(x: torch.Tensor, blk: int, mean: torch.Tensor | None = None, rows_per_head: int | None = None):
    """x: [..., D] fp16 (flattened to [N, D]) -> (int8 [N, D], fp16 scales [N/blk]).
    mean: optional [B,H,1,D] fp16 subtracted per head before quantisation (needs rows_per_head = S)."""
    _need_cuda(x, mean)
    assert x.dtype == torch.float16
    D = x.shape[-1]
    x2 = x.contiguous().view(-1, D)
    N = x2.shape[0]
    out = torch.empty((N, D), dtype=torch.int8, device=x.device)
    scales = torch.empty((N // blk,), dtype=torch.float16, device=x.device)
    if mean is not None:
        assert mean.dtype == torch.float16 and rows_per_head is not None
        mean = mean.contiguous()
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.qa_quant_block(_lib.ptr(x2), _lib.ptr(mean), _lib.ptr(out), _lib.ptr(scales), N, D, blk,
                                    rows_per_head or N, _lib.cur_stream()), "qa_quant_block")
    return out, scales


def int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, Sq, Sk, D, Bq=128, Bkv=128, nsplit=2, want_lse32=True,
                      ring_state=False):
    """Fused int8 attention forward over pre-quantised operands (qa_int8_fwd).
    Returns (O fp16 [BH*Sq, D], lse16 [BH*Sq], lse32 or None); with ring_state=True returns the unnormalised
    (O_acc fp32 [BH*Sq, D], m fp32 [BH*Sq], l fp32 [BH*Sq]) of this K/V shard instead."""
    _need_cuda(q_i8, k_i8, v_i8, sq, sk, sv)
    dev = q_i8.device
    L = _lib.lib()
    if ring_state:
        o_acc = torch.empty((BH * Sq, D), dtype=torch.float32, device=dev)
        m = torch.empty((BH * Sq,), dtype=torch.float32, device=dev)
        l = torch.empty((BH * Sq,), dtype=torch.float32, device=dev)
        O = lse16 = lse32 = None
    else:
        O = torch.empty((BH * Sq, D), dtype=torch.float16, device=dev)
        lse16 = torch.empty((BH * Sq,), dtype=torch.float16, device=dev)
        lse32 = torch.empty((BH * Sq,), dtype=torch.float32, device=dev) if want_lse32 else None
        o_acc = m = l = None
    with torch.cuda.device(dev):
        _lib.check(L.qa_int8_fwd(_lib.ptr(q_i8), _lib.ptr(k_i8), _lib.ptr(v_i8), _lib.ptr(sq), _lib.ptr(sk), _lib.ptr(sv),
                                 _lib.ptr(O), _lib.ptr(lse16), _lib.ptr(lse32), _lib.ptr(o_acc), _lib.ptr(m), _lib.ptr(l),
                                 BH, Sq, Sk, D, Bq, Bkv, nsplit, _lib.cur_stream()), "qa_int8_fwd")
    if ring_state:
        return o_acc, m, l
    return O, lse16, lse32
