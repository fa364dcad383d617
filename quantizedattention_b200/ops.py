"""Low-level Python bindings of the C-ABI kernels (one function per `qa_*` entry point).

Torch owns every buffer (inputs, outputs, workspaces are `tensor.data_ptr()`); the library never
allocates device memory and never synchronises.  Everything runs on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib


# Optional kernel timing hook: set `TIMING = []` and every hot kernel launch appends (name, start_evt, end_evt)
# recorded on the launching (current) stream; bench.py uses it for the live roofline measurement.
TIMING = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if TIMING is not None:
            self.b.record()
            TIMING.append((self.name, self.a, self.b))
        return False


def _unwrap(t):
    """Strip functorch wrappers (torch.func.jvp hands GradTrackingTensors to Function.jvp) to reach the storage."""
    F = torch._C._functorch
    while t is not None and (F.is_gradtrackingtensor(t) or F.is_batchedtensor(t) or F.is_functionaltensor(t)):
        t = F.get_unwrapped(t)
    return t


class _raw_mode:
    """Run raw-pointer kernels below any active functorch transform (inside Function.jvp under torch.func.jvp every
    torch op would otherwise re-wrap its outputs)."""

    def __enter__(self):
        from torch._functorch.pyfunctorch import temporarily_clear_interpreter_stack
        self.cm = temporarily_clear_interpreter_stack()
        return self.cm.__enter__()

    def __exit__(self, *exc):
        return self.cm.__exit__(*exc)


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


ROUNDING = {"trunc": 0, "nearest": 1}       # int8 rounding of the quantisers: the reference truncates (LEDGER I-3)
QUANT_MODES = {"trunc": 0, "nearest": 1, "e4m3": 2}   # qa_quant_block `rounding`: 2 = fp8 e4m3 codes (scale = amax / 448)
FLAG_NEAREST, FLAG_CAUSAL, FLAG_BWD_8WARP = 1, 2, 4      # `flags` of qa_int8_fwd / qa_int8_bwd (include/qattn.h)


def _flags(rounding: str, causal: bool) -> int:
    return (FLAG_NEAREST if ROUNDING[rounding] else 0) | (FLAG_CAUSAL if causal else 0)


def quant_block(x: torch.Tensor, blk: int, mean: torch.Tensor | None = None, rows_per_head: int | None = None,
                rounding: str = "trunc"):
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
                                    rows_per_head or N, QUANT_MODES[rounding], _lib.cur_stream()), "qa_quant_block")
    return out, scales


def k_token_sum(k: torch.Tensor) -> torch.Tensor:
    """fp32 token sums [B,H,1,D] of a (sequence-shard of) K (qa_k_token_sum)."""
    _need_cuda(k)
    assert k.dtype == torch.float16 and k.dim() == 4
    k = k.contiguous()
    B, H, S, D = k.shape
    L = _lib.lib()
    ws = torch.empty(L.qa_k_mean_workspace_bytes(B, H, S, D), dtype=torch.uint8, device=k.device)
    out = torch.empty((B, H, 1, D), dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        _lib.check(L.qa_k_token_sum(_lib.ptr(k), _lib.ptr(out), _lib.ptr(ws), ws.numel(), B, H, S, D, _lib.cur_stream()),
                   "qa_k_token_sum")
    return out


def int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, Sq, Sk, D, Bq=128, Bkv=128, nsplit=0, want_lse32=True,
                      ring_state=False, state_in=None, rounding: str = "trunc", causal: bool = False, sk_valid: int | None = None):
    """Fused int8 attention forward over pre-quantised operands (qa_int8_fwd).
    nsplit: 0 = the default kernel for the tile (Bkv = 128: two-stage softmax with magic accumulators); 1 = single-stage
    softmax (one thread per row); 2 = two-stage softmax without magic accumulators (kept for A/B comparison).
    Returns (O fp16 [BH*Sq, D], lse16 [BH*Sq], lse32 or None); with ring_state=True returns the unnormalised
    (O_acc fp32 [BH*Sq, D], m fp32 [BH*Sq], l fp32 [BH*Sq]) of this K/V shard instead.
    sk_valid: ragged sequence zero-padded per head to Sk: keys >= sk_valid have weight 0 (qa_int8_fwd_ragged)."""
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
    si = state_in if state_in is not None else (None, None, None)       # (o_acc, m, l) of the earlier K/V shards
    with torch.cuda.device(dev), _timed("int8_fwd"):
        _lib.check(L.qa_int8_fwd_ragged(_lib.ptr(q_i8), _lib.ptr(k_i8), _lib.ptr(v_i8), _lib.ptr(sq), _lib.ptr(sk),
                                        _lib.ptr(sv), _lib.ptr(O), _lib.ptr(lse16), _lib.ptr(lse32), _lib.ptr(o_acc),
                                        _lib.ptr(m), _lib.ptr(l), _lib.ptr(si[0]), _lib.ptr(si[1]), _lib.ptr(si[2]),
                                        BH, Sq, Sk, Sk if sk_valid is None else int(sk_valid), D, Bq, Bkv, nsplit,
                                        _flags(rounding, causal), _lib.cur_stream()), "qa_int8_fwd")
    if ring_state:
        return o_acc, m, l
    return O, lse16, lse32


def bwd_delta(dO: torch.Tensor, O: torch.Tensor, want_bf16_copy: bool = False):
    """delta = rowsum(dO * O) in fp32 (qa_bwd_delta).  fp16 inputs (int8 path) or fp32 inputs (bf16 path, where a
    bf16 copy of dO can be emitted in the same pass)."""
    _need_cuda(dO, O)
    D = dO.shape[-1]
    dO2, O2 = dO.contiguous().view(-1, D), O.contiguous().view(-1, D)
    assert dO2.dtype == O2.dtype and dO2.dtype in (torch.float16, torch.float32)
    n = dO2.shape[0]
    delta = torch.empty((n,), dtype=torch.float32, device=dO.device)
    copy = torch.empty((n, D), dtype=torch.bfloat16, device=dO.device) if want_bf16_copy else None
    L = _lib.lib()
    with torch.cuda.device(dO.device):
        _lib.check(L.qa_bwd_delta(_lib.ptr(dO2), _lib.ptr(O2), _lib.ptr(delta), _lib.ptr(copy), n, D,
                                  0 if dO2.dtype == torch.float16 else 1, _lib.cur_stream()), "qa_bwd_delta")
    return (delta, copy) if want_bf16_copy else delta


def cast_f32(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 -> fp16 / bf16 (qa_cast_f32)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and dtype in (torch.float16, torch.bfloat16)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.qa_cast_f32(_lib.ptr(x), _lib.ptr(out), x.numel(), 0 if dtype == torch.float16 else 1,
                                 _lib.cur_stream()), "qa_cast_f32")
    return out


def int8_bwd_prequant(q_i8, k_i8, v_i8, do_i8, sq, sk, sv, s_do, lse32, delta, k_mean, BH, S, D, Bq=128, Bkv=128,
                      rounding: str = "trunc", causal: bool = False, kernel: str = "ws", s_valid: int | None = None):
    """Fused int8 backward over pre-quantised operands (qa_int8_bwd).  Returns (dq, dk, dv) fp16 [BH*S, D].
    kernel: "ws" = the warp-specialised kernel (default), "8warp" = the 8-warp kernel (QA_FLAG_BWD_8WARP)."""
    if kernel not in ("ws", "8warp"):
        raise ValueError('kernel must be "ws" or "8warp"')
    flags = _flags(rounding, causal) | (FLAG_BWD_8WARP if kernel == "8warp" else 0)
    _need_cuda(q_i8, k_i8, v_i8, do_i8)
    dev = q_i8.device
    dq_ws = torch.zeros((BH * S, D), dtype=torch.float32, device=dev)
    rowsum_ws = torch.zeros((BH * S,), dtype=torch.float32, device=dev) if k_mean is not None else None
    alloc = torch.empty if s_valid is None else torch.zeros     # ragged: k-tiles without a valid key are not written
    dk = alloc((BH * S, D), dtype=torch.float16, device=dev)
    dv = alloc((BH * S, D), dtype=torch.float16, device=dev)
    dq = torch.empty((BH * S, D), dtype=torch.float16, device=dev)
    if k_mean is not None:
        k_mean = k_mean.contiguous()
    L = _lib.lib()
    with torch.cuda.device(dev):
        with _timed("int8_bwd"):
            _lib.check(L.qa_int8_bwd_ragged(_lib.ptr(q_i8), _lib.ptr(k_i8), _lib.ptr(v_i8), _lib.ptr(do_i8), _lib.ptr(sq), _lib.ptr(sk),
                                            _lib.ptr(sv), _lib.ptr(s_do), _lib.ptr(lse32), _lib.ptr(delta), _lib.ptr(rowsum_ws),
                                            _lib.ptr(dq_ws), _lib.ptr(dk), _lib.ptr(dv), BH, S, S if s_valid is None else int(s_valid),
                                            D, Bq, Bkv, flags, _lib.cur_stream()), "qa_int8_bwd")
        _lib.check(L.qa_int8_bwd_finalize(_lib.ptr(dq_ws), _lib.ptr(rowsum_ws), _lib.ptr(k_mean), _lib.ptr(dq), BH, S, D,
                                          _lib.cur_stream()), "qa_int8_bwd_finalize")
    return dq, dk, dv


def int8_bwd_sage(q_i8, k_i8, v_fp16, do_i8, dO_fp16, sq, sk, s_do, lse32, delta, k_mean, BH, S, D, Bq=128, Bkv=128,
                  rounding: str = "trunc"):
    """SageBwd backward (qa_int8_bwd_sage): dP = dO V^T in fp16 from the unquantised tensors, everything else int8.
    v_fp16, dO_fp16: [BH*S, D] fp16.  Returns (dq, dk, dv) fp16 [BH*S, D]."""
    _need_cuda(q_i8, k_i8, v_fp16, do_i8, dO_fp16)
    assert v_fp16.dtype == torch.float16 and dO_fp16.dtype == torch.float16
    dev = q_i8.device
    dq_ws = torch.zeros((BH * S, D), dtype=torch.float32, device=dev)
    rowsum_ws = torch.zeros((BH * S,), dtype=torch.float32, device=dev) if k_mean is not None else None
    dk = torch.empty((BH * S, D), dtype=torch.float16, device=dev)
    dv = torch.empty((BH * S, D), dtype=torch.float16, device=dev)
    dq = torch.empty((BH * S, D), dtype=torch.float16, device=dev)
    if k_mean is not None:
        k_mean = k_mean.contiguous()
    L = _lib.lib()
    with torch.cuda.device(dev):
        with _timed("int8_bwd_sage"):
            _lib.check(L.qa_int8_bwd_sage(_lib.ptr(q_i8), _lib.ptr(k_i8), _lib.ptr(v_fp16.contiguous()), _lib.ptr(do_i8),
                                          _lib.ptr(dO_fp16.contiguous()), _lib.ptr(sq), _lib.ptr(sk), _lib.ptr(s_do), _lib.ptr(lse32),
                                          _lib.ptr(delta), _lib.ptr(rowsum_ws), _lib.ptr(dq_ws), _lib.ptr(dk), _lib.ptr(dv), BH, S, D,
                                          Bq, Bkv, _flags(rounding, False), _lib.cur_stream()), "qa_int8_bwd_sage")
        _lib.check(L.qa_int8_bwd_finalize(_lib.ptr(dq_ws), _lib.ptr(rowsum_ws), _lib.ptr(k_mean), _lib.ptr(dq), BH, S, D,
                                          _lib.cur_stream()), "qa_int8_bwd_finalize")
    return dq, dk, dv


def bf16_fwd_key_step(Sq: int, nsplit: int = 0) -> int:
    """Keys per online-softmax step of the kernels qa_bf16_fwd dispatches to (128 for both since the two-tile kernel moved
    from 64-key to 128-key steps: a tcgen05.mma with N = 64 costs as much as one with N = 128, profiles/r02_mma_rate.txt): the running maximum (and its bias
    correction) advances once per step, so an oracle comparison at rounding level must use the same `tile_k`.
    The reference's own default is 32 (attention_bf16.py:139); the choice is mathematically neutral."""
    return 128


def _pad_seq(t, Sp: int, value: float = 0.0):
    """[B,H,S,D] (or [B*H,S]) -> zero-padded along the token axis to Sp (ragged sequences: the kernels mask keys >= S)."""
    dim = 2 if t.dim() == 4 else 1
    S = t.shape[dim]
    if S == Sp:
        return t.contiguous()
    shape = list(t.shape)
    shape[dim] = Sp
    out = torch.full(shape, value, dtype=t.dtype, device=t.device)
    out.narrow(dim, 0, S).copy_(t)
    return out


def _ceil128(S: int) -> int:
    return (S + 127) // 128 * 128


BF16_RESCALE_TAU = 8.0      # default lazy-rescale threshold of qa_bf16_fwd (log2 units); 0 = the reference's step-by-step maximum


def bf16_fwd(q, k, v, causal: bool, nsplit: int = 0, rescale_tau: float | None = None):
    """Bias-corrected bf16 flash attention forward (qa_bf16_fwd_ragged).  q,k fp16, v bf16 [B,H,S,D] ->
    (O fp32 [B,H,Sq,D], lse fp32 [B*H,Sq]).  Sequence lengths that are not multiples of 128 (the reference's clamped last
    hl.tile) are zero-padded here and masked in the kernel."""
    _need_cuda(q, k, v)
    B, H, Sq, D = q.shape
    Sk = k.shape[2]
    Sqp, Skp = _ceil128(Sq), _ceil128(Sk)
    q, k, v = _pad_seq(q, Sqp), _pad_seq(k, Skp), _pad_seq(v, Skp)
    O = torch.empty((B, H, Sqp, D), dtype=torch.float32, device=q.device)
    lse = torch.empty((B * H, Sqp), dtype=torch.float32, device=q.device)
    L = _lib.lib()
    with torch.cuda.device(q.device), _timed("bf16_fwd"):
        rc = L.qa_bf16_fwd_ragged(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(O), _lib.ptr(lse), B * H, Sqp, Skp, Sk, D,
                                  1 if causal else 0, nsplit, float(BF16_RESCALE_TAU if rescale_tau is None else rescale_tau),
                                  _lib.cur_stream())
        _lib.check(rc, "qa_bf16_fwd")
    if Sqp != Sq:
        O, lse = O[:, :, :Sq].contiguous(), lse[:, :Sq].contiguous()
    return O, lse


def jvp_fwd(q, k, v, tq, tk, tv, nsplit: int = 2):
    """Forward-mode JVP attention (qa_jvp_fwd_ragged).  Six fp32 [B,H,S,D] tensors -> (O, tO fp32 [B,H,Sq,D], lse [B*H,Sq]).
    Operands are rounded to bf16 for the tensor cores (fp32 accumulation).  Sequence lengths that are not
    multiples of 128 are zero-padded here and masked in the kernel."""
    q, k, v, tq, tk, tv = [_unwrap(t) for t in (q, k, v, tq, tk, tv)]
    _need_cuda(q, k, v, tq, tk, tv)
    with _raw_mode():
        B, H, Sq, D = q.shape
        Sk = k.shape[2]
        Sqp, Skp = _ceil128(Sq), _ceil128(Sk)
        b16 = [cast_f32(t, torch.bfloat16) if t.dtype == torch.float32 else t.to(torch.bfloat16).contiguous()
               for t in (q, tq, k, tk, v, tv)]
        b16 = [_pad_seq(t, Sqp if i < 2 else Skp) for i, t in enumerate(b16)]
        O = torch.empty((B, H, Sqp, D), dtype=torch.float32, device=q.device)
        tO = torch.empty_like(O)
        lse = torch.empty((B * H, Sqp), dtype=torch.float32, device=q.device)
        L = _lib.lib()
        with torch.cuda.device(q.device), _timed("jvp_fwd"):
            _lib.check(L.qa_jvp_fwd_ragged(*[_lib.ptr(t) for t in b16], _lib.ptr(O), _lib.ptr(tO), _lib.ptr(lse), B * H, Sqp, Skp,
                                           Sk, D, nsplit, _lib.cur_stream()), "qa_jvp_fwd")
        if Sqp != Sq:
            O, tO, lse = O[:, :, :Sq].contiguous(), tO[:, :, :Sq].contiguous(), lse[:, :Sq].contiguous()
    return O, tO, lse


def bf16_bwd(q, k, v, O, lse, causal: bool, dO, variant: int = 0):
    """Recompute backward of the bf16 path (qa_bwd_delta + qa_bf16_bwd_ragged).  q,k fp16; v bf16; O, dO fp32 [B,H,S,D];
    lse fp32 [B*H,S].  Returns fp32 (dq, dk, dv) [B,H,S,D].  variant: 0 = default kernels, 1 = phase-sequential kernel.
    Ragged S: every operand is zero-padded to a multiple of 128, lse with a large value (P = 0 for the padded query rows)."""
    _need_cuda(q, k, v, O, lse, dO)
    B, H, S, D = q.shape
    assert k.shape[2] == S, "backward is self-attention only (LEDGER I-11)"
    Sp = _ceil128(S)
    q, k, v = _pad_seq(q, Sp), _pad_seq(k, Sp), _pad_seq(v, Sp)
    dO = _pad_seq(dO.to(torch.float32), Sp)
    O = _pad_seq(O.to(torch.float32), Sp)
    lse = _pad_seq(lse.to(torch.float32).view(B * H, S), Sp, 1.0e30)
    delta, dO_bf16 = bwd_delta(dO, O, want_bf16_copy=True)
    dq = torch.zeros((B, H, Sp, D), dtype=torch.float32, device=q.device)
    dk = torch.empty_like(dq)
    dv = torch.empty_like(dq)
    L = _lib.lib()
    with torch.cuda.device(q.device), _timed("bf16_bwd"):
        _lib.check(L.qa_bf16_bwd_ragged(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(dO_bf16), _lib.ptr(dO), _lib.ptr(lse),
                                        _lib.ptr(delta), _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv), B * H, Sp, S, D,
                                        1 if causal else 0, int(variant), _lib.cur_stream()), "qa_bf16_bwd")
    if Sp != S:
        dq, dk, dv = [t[:, :, :S].contiguous() for t in (dq, dk, dv)]
    return dq, dk, dv
