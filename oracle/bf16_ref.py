"""Eager restatement of the bf16 bias-corrected flash attention path (attention_bf16.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  `literal` modes are pinned bit-for-bit against
the unmodified reference by tests/test_oracle_golden.py.

Modes:
  literal  -- the source: logits fp32->fp16->bf16, strict causal with a finite -126 fill applied
              BEFORE scaling (leaky), near-max predicate comparing unscaled S with scaled m,
              backward with dS = S*(dP - D) and qk_scale on dQ/dK.
  contract -- what the CUDA kernels implement: scaled logits u = bf16(S*qk_scale) rounded ONCE from the fp32
              accumulator (the literal rounds fp32->fp16->bf16, scales, rounds again; LEDGER B-9), masked weight
              exactly 0 for rows >= 1 (fully-masked tiles skipped), row 0 = uniform over ALL keys
              (what both the reference kernel and its baseline produce), the near-max predicate
              evaluated in ONE (scaled) domain, backward with dS = P*(dP - delta) and sm_scale.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504
BETA = 2.0            # attention_bf16.py:146


def _bias_corrected_max(m_prev, u_or_s, cmp_vals):
    """attention_bf16.py:236-264.  `u_or_s`: values whose row max (already scaled, bf16) feeds
    the running max; `cmp_vals`: values compared against (m' - 1e-3)."""
    m_new = torch.max(m_prev, u_or_s).to(torch.bfloat16)
    approx = cmp_vals >= (m_new - 1e-3)
    many = torch.sum(approx, dim=-1, keepdim=True) > 1
    m_new = torch.where(many & (m_new > 0), BETA * m_new, m_new)
    zero = torch.tensor(0.0, dtype=torch.bfloat16)
    m_new = torch.where(many & (m_new < 0), zero, m_new)
    return m_new


def bf16_fwd(q, k, v, causal: bool, tile_k: int = 32, mode: str = "literal", lazy_tau: float = 0.0):
    """helion_atten_bf16_fwd_training (attention_bf16.py:195-294), vectorised over q rows
    (rows are independent; the `begin_q < end_k` guard only skips a no-op mask).
    q,k fp16, v bf16 [B,H,S,D] -> (O fp32 [B,H,S,D], lse fp32 [B*H,S]).  tile_k = k-tile width
    (the literal result depends on it through the running bf16 max).  lazy_tau (contract mode only): a new running
    maximum is adopted only when it exceeds the current one by more than lazy_tau log2 units (the CUDA kernel's
    lazy rescale, include/qattn.h qa_bf16_fwd_ex); 0 = the reference's step-by-step maximum."""
    assert mode in ("literal", "contract")
    B, H, S, D = q.shape
    Sk = k.shape[2]
    G = B * H
    qg = q.reshape(G, S, D)
    kgT = k.reshape(G, Sk, D).transpose(1, 2)
    vg = v.reshape(G, Sk, D)
    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E

    m = torch.full((G, S, 1), float("-inf"), dtype=torch.bfloat16)      # :197
    l = torch.full((G, S, 1), 1.0, dtype=torch.float32)                 # :198
    O = torch.zeros((G, S, D), dtype=torch.float32)                     # :199
    qi = torch.arange(S)[:, None]

    for kb in range(0, Sk, tile_k):
        ke = min(kb + tile_k, Sk)
        if mode == "literal":
            S16 = torch.bmm(qg, kgT[:, :, kb:ke])                        # :215 fp16 result
            Sb = S16.to(torch.bfloat16)                                  # :216
        else:
            Sf = torch.bmm(qg.float(), kgT[:, :, kb:ke].float())          # fp32 logits (tensor-core accumulator)
            Sb = (Sf * qk_scale).to(torch.bfloat16)                        # contract: ONE rounding, u = bf16(S * qk_scale)
        if causal:
            keep = (qi - torch.arange(kb, ke)[None, :]) > 0              # :226 strict
            if mode == "literal":
                Sb = torch.where(keep, Sb, torch.tensor([-126], dtype=torch.bfloat16))   # :228-233
            else:
                Sb = torch.where(keep, Sb, torch.tensor([float("-inf")], dtype=torch.bfloat16))
        if mode == "literal":
            m_new = _bias_corrected_max(m, torch.amax(Sb, -1, keepdim=True) * qk_scale, Sb)  # :236-264
            Ssh = Sb * qk_scale - m_new                                  # :267
        else:
            u = Sb                                                       # bf16 scaled logits (already scaled above)
            m_new = _bias_corrected_max(m, torch.amax(u, -1, keepdim=True), u)
            if lazy_tau > 0:
                adopt = (m_new - m).float() > lazy_tau                   # bf16 difference, as the kernel computes it
                m_new = torch.where(adopt, m_new, m)
            # rows whose every key so far is masked keep m = -inf; avoid (-inf) - (-inf)
            m_fin = torch.where(torch.isinf(m_new), torch.zeros_like(m_new), m_new)
            Ssh = u - m_fin
        P = torch.exp2(Ssh.to(torch.float32)).to(torch.bfloat16)         # :269
        l_new = torch.sum(P.to(torch.float32), -1, keepdim=True)         # :274
        resc = torch.exp2((m - m_new).to(torch.float32)).to(torch.bfloat16)   # :276
        if mode == "contract":
            resc = torch.where(torch.isinf(m_new), torch.zeros_like(resc), resc)
        m = m_new
        l = l * resc.to(torch.float32) + l_new                           # :279
        O = O * resc                                                     # :280
        O = O + torch.bmm(P.float(), vg[:, kb:ke].float())               # :285 (bf16 MMA, fp32 acc)

    lse = m.squeeze(-1) + torch.log2(l).squeeze(-1)                      # :288 (bf16 + fp32 -> fp32)
    Ofin = O / l                                                         # :293
    if mode == "contract" and causal:
        # row 0 has no visible key: reference kernel and baseline both give the uniform average
        # over all keys (LEDGER B-1).  lse uses the backward's -128 fill (attention_bf16.py:384).
        Ofin[:, 0] = vg.float().mean(dim=1)
        lse[:, 0] = -128.0 + math.log2(Sk)
    return Ofin.view(B, H, S, D), lse


def bf16_bwd(q, k, v, O, lse, causal: bool, dO, mode: str = "literal", tile_q: int = 16, tile_k: int = 16):
    """helion_flash_atten_2_algo_4_bwd (attention_bf16.py:361-444).

    literal: fp32 everywhere, dS = S*(dP-D) (B-5), qk_scale on dQ/dK (B-6), -128 fill (B-4);
             result is tile-size independent except fp32 summation order, so the restatement
             loops over tiles exactly like the source to stay bit-exact.
    contract: P = exp2(S - lse) with masked weight 0 (row 0: uniform over all keys, no dS),
             dS = P*(dP - delta), sm_scale; fp32 math (the CUDA kernel rounds MMA operands to
             fp16/bf16 -- tolerance-level parity)."""
    assert mode in ("literal", "contract")
    B, H, S, D = q.shape
    G = B * H
    qf = q.to(torch.float32).reshape(G, S, D)
    kf = k.to(torch.float32).reshape(G, S, D)
    vf = v.to(torch.float32).reshape(G, S, D)
    Of = O.reshape(G, S, D)
    dOf = dO.reshape(G, S, D)
    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E
    if mode == "contract":
        Sx = qk_scale * torch.bmm(qf, kf.transpose(1, 2))
        P = torch.exp2(Sx - lse[:, :, None])
        if causal:
            keep = (torch.arange(S)[:, None] - torch.arange(S)[None, :]) > 0
            P = torch.where(keep[None], P, torch.zeros(()))
            P[:, 0, :] = 1.0 / S
        dv = torch.bmm(P.transpose(1, 2), dOf)
        dP = torch.bmm(dOf, vf.transpose(1, 2))
        delta = torch.sum(dOf * Of, dim=-1, keepdim=True)
        dS = P * (dP - delta)
        if causal:
            dS[:, 0, :] = 0.0
        dq = sm_scale * torch.bmm(dS, kf)
        dk = sm_scale * torch.bmm(dS.transpose(1, 2), qf)
        return dq.view(B, H, S, D), dk.view(B, H, S, D), dv.view(B, H, S, D)

    dq = torch.zeros_like(qf)
    dk = torch.zeros((G, S, D))
    dv = torch.zeros_like(vf)
    kT = kf.transpose(-1, -2)
    for kb in range(0, S, tile_k):
        ke = min(kb + tile_k, S)
        kt = kT[:, :, kb:ke]
        vt = vf[:, kb:ke]
        dk_t = torch.zeros((G, ke - kb, D))
        dv_t = torch.zeros((G, ke - kb, D))
        for qb in range(0, S, tile_q):
            qe = min(qb + tile_q, S)
            qt = qf[:, qb:qe]
            Sx = qk_scale * torch.bmm(qt, kt)                            # :376-377
            if causal and qb < ke:
                keep = (torch.arange(qb, qe)[:, None] - torch.arange(kb, ke)[None, :]) > 0
                Sx = torch.where(keep, Sx, torch.tensor([-128]))        # :384-389
            P = torch.exp2(Sx - lse[:, qb:qe, None])                     # :392
            dOt = dOf[:, qb:qe]
            dv_t = dv_t + torch.bmm(P.transpose(1, 2), dOt)              # :399
            dP = torch.bmm(dOt, vt.transpose(1, 2))                      # :405
            Dv = torch.sum(dOt * Of[:, qb:qe], dim=-1, keepdim=True)     # :416
            dS = Sx * (dP - Dv)                                          # :421
            dq[:, qb:qe] = dq[:, qb:qe] + torch.bmm(qk_scale * dS, kt.transpose(-1, -2))   # :428-432
            dk_t = dk_t + torch.bmm(qk_scale * dS.transpose(-1, -2), qt)                    # :436-441
        dk[:, kb:ke] = dk_t
        dv[:, kb:ke] = dv_t
    return dq.view(B, H, S, D), dk.view(B, H, S, D), dv.view(B, H, S, D)
