"""Host-staged int8 attention: q, k, v (and dO) live in pinned HOST memory; results come back to pinned host memory.

The batch x head axis is the path's natural shard (every (b, h) is an independent attention problem, the
K token mean and the quantisation blocks never straddle heads), so a head chunk is a complete unit of work.  Three CUDA
streams run the chunks as a pipeline,

    copy-in stream  :  H2D(q,k,v,dO of chunk i+1)
    compute stream  :  sage_attention_3_int8 forward + backward of chunk i   (attention_int8.py:434-451, :20-95)
    copy-out stream :  D2H(O,dq,dk,dv of chunk i-1)

with a few device staging slots for the inputs and CUDA events between the stages.  PCIe is full duplex, so the whole
job costs about max(H2D, compute, D2H) instead of their sum.  This is the call `bench.py` times as `e2e`.
"""
from __future__ import annotations

import torch

from . import attention_int8 as A


class HostStagedSageAttention:
    """Reusable pipeline object (device staging buffers and streams are allocated once per shape)."""

    def __init__(self, device=None, heads_per_chunk: int = 16, slots: int = 3):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.heads_per_chunk = heads_per_chunk
        self.slots = max(2, slots)                                 # device staging slots for the inputs
        self.trace = None                                          # set to [] to collect (stage, chunk, start, end) timing events
        self._key = None

    def _span(self, stage, chunk, stream):
        """When tracing, record the start of a (stage, chunk) span on `stream` and return its end event."""
        if self.trace is None:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        self.trace.append((stage, chunk, a, b))
        return b

    def _setup(self, BH, S, D, with_grad):
        hc = min(self.heads_per_chunk, BH)
        while BH % hc:
            hc -= 1
        key = (BH, S, D, hc, with_grad, self.slots)
        if key == self._key:
            return
        dev = self.device
        n_in = 4 if with_grad else 3
        self.hc = hc
        # chunk schedule: the first H2D and the last D2H cannot overlap anything, and the D2H lane runs one chunk (+ its compute)
        # behind the H2D lane, so the edge chunks taper geometrically (hc/8, hc/8, hc/4, hc/2, hc, ..., hc, hc/2, hc/4, hc/8, hc/8)
        # when there are enough heads for that: the un-overlapped head and tail of a step shrink to the copies of hc/8 heads
        if hc % 8 == 0 and BH >= 4 * hc:
            up = [hc // 8, hc // 8, hc // 4, hc // 2]
            sizes = up + [hc] * (BH // hc - 2) + up[::-1]
        elif hc % 4 == 0 and BH >= 4 * hc:
            sizes = [hc // 4, 3 * hc // 4] + [hc] * (BH // hc - 2) + [3 * hc // 4, hc // 4]
        else:
            sizes = [hc] * (BH // hc)
        self.sched, h0 = [], 0
        for nh in sizes:
            self.sched.append((h0, nh))
            h0 += nh
        assert h0 == BH
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.inbuf = [[torch.empty((1, hc, S, D), dtype=torch.float16, device=dev) for _ in range(n_in)] for _ in range(self.slots)]
        # results are copied into persistent staging slots on the compute stream, so every temporary of the attention
        # call is allocated and freed on ONE stream (the caching allocator then never has to cudaMalloc, which would
        # synchronise the device and stall the pipeline) and the copy-out stream only ever reads these buffers
        self.outbuf = [[torch.empty((hc, S, D), dtype=torch.float16, device=dev) for _ in range(4 if with_grad else 1)]
                       for _ in range(self.slots)]
        self._key = key

    @staticmethod
    def _check_host(*ts):
        for t in ts:
            if t is None:
                continue
            if t.is_cuda or t.dtype != torch.float16 or not t.is_contiguous():
                raise TypeError("host-staged attention takes contiguous fp16 HOST tensors [B,H,S,D]")
            if not t.is_pinned():
                raise RuntimeError("host tensors must be pinned (tensor.pin_memory()): pageable memory cannot overlap copies")

    def __call__(self, q, k, v, dO=None, out=None):
        """q, k, v (and optionally dO): pinned fp16 host tensors [B,H,S,D].
        Returns O (and dq, dk, dv when dO is given) as pinned fp16 host tensors; `out` supplies them if given."""
        self._check_host(q, k, v, dO)
        B, H, S, D = q.shape
        if k.shape != q.shape or v.shape != q.shape or (dO is not None and dO.shape != q.shape):
            raise ValueError("q, k, v, dO must share one [B,H,S,D] shape (self-attention)")
        BH = B * H
        with_grad = dO is not None
        self._setup(BH, S, D, with_grad)
        hc, dev = self.hc, self.device
        n_out = 4 if with_grad else 1
        if out is None:
            out = [torch.empty((B, H, S, D), dtype=torch.float16).pin_memory() for _ in range(n_out)]
        self._check_host(*out)
        srcs = [t.view(BH, S, D) for t in ((q, k, v, dO) if with_grad else (q, k, v))]
        dsts = [t.view(BH, S, D) for t in out]
        sched = self.sched
        n = len(sched)
        cur = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(cur)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)                                    # the pipeline starts after the caller's prior work
        ev_in, ev_cmp, ev_out = [None] * n, [None] * n, [None] * n
        for i in range(n + 2):
            if i < n:                                              # ---- stage 1: H2D of chunk i
                slot = i % self.slots
                with torch.cuda.stream(self.s_in):
                    if i >= self.slots:
                        self.s_in.wait_event(ev_cmp[i - self.slots])   # the slot's previous reader has finished
                    sp = self._span("h2d", i, self.s_in)
                    h0, nh = sched[i]
                    for dst, src in zip(self.inbuf[slot], srcs):
                        dst.view(hc, S, D)[:nh].copy_(src[h0:h0 + nh], non_blocking=True)
                    if sp is not None:
                        sp.record(self.s_in)
                    ev_in[i] = torch.cuda.Event()
                    ev_in[i].record(self.s_in)
            j = i - 1
            if 0 <= j < n:                                         # ---- stage 2: forward (+ backward) of chunk j
                slot = j % self.slots
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(ev_in[j])
                    sp = self._span("compute", j, self.s_cmp)
                    nh = sched[j][1]
                    bufs = [t[:, :nh] for t in self.inbuf[slot]]
                    if j >= self.slots:
                        self.s_cmp.wait_event(ev_out[j - self.slots])   # the output slot has been copied out
                    if with_grad:
                        qr, kr, vr = (t.detach().requires_grad_() for t in bufs[:3])
                        O = A.sage_attention_3_int8(qr, kr, vr)
                        O.backward(bufs[3])
                        res = (O.detach(), qr.grad, kr.grad, vr.grad)
                    else:
                        with torch.no_grad():
                            res = (A.sage_attention_3_int8(*bufs[:3]),)
                    for dst, src in zip(self.outbuf[slot], res):
                        dst[:nh].copy_(src.view(nh, S, D))
                    del res
                    if sp is not None:
                        sp.record(self.s_cmp)
                    ev_cmp[j] = torch.cuda.Event()
                    ev_cmp[j].record(self.s_cmp)
            m = i - 2
            if 0 <= m < n:                                         # ---- stage 3: D2H of chunk m
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_cmp[m])
                    sp = self._span("d2h", m, self.s_out)
                    h0, nh = sched[m]
                    for dst, src in zip(dsts, self.outbuf[m % self.slots]):
                        dst[h0:h0 + nh].copy_(src[:nh], non_blocking=True)
                    if sp is not None:
                        sp.record(self.s_out)
                    ev_out[m] = torch.cuda.Event()
                    ev_out[m].record(self.s_out)
        cur.wait_event(ev_out[n - 1])                              # results are complete for work queued after the call
        cur.wait_event(ev_cmp[n - 1])
        return out[0] if not with_grad else tuple(out)


_DEFAULT = {}


def sage_attention_3_int8_host(q, k, v, dO=None, out=None, heads_per_chunk: int = 16, device=None):
    """Functional form of HostStagedSageAttention (one cached pipeline per device and chunk size).
    Returns O, or (O, dq, dk, dv) when dO is given; the copies are asynchronous on the current stream's timeline:
    synchronise (or record an event) before reading the host results."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev, heads_per_chunk)
    if key not in _DEFAULT:
        _DEFAULT[key] = HostStagedSageAttention(dev, heads_per_chunk)
    return _DEFAULT[key](q, k, v, dO, out)
