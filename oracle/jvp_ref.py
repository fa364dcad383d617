"""Eager restatement of the forward-mode JVP attention (attention_jvp.py:129-190).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Pinned bit-for-bit against the unmodified
reference by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504


def jvp_fwd(q, k, v, tq, tk, tv, tile_k: int = 16, operand_dtype=None):
    """helion_attention_jvp_forward_fp32: six contractions per k-tile, fp32 state.
    Returns (O, tO [B,H,S,D] fp32, lse [B*H,S] fp32).

    operand_dtype=torch.bfloat16 emulates the CUDA kernel's MMA operand rounding (q,k,v,tq,tk,tv,
    P and H rounded to bf16 before each contraction; fp32 accumulate) -- LEDGER J-2."""
    B, H, S, D = q.shape
    Sk = k.shape[2]
    G = B * H
    r_ = (lambda t: t.to(operand_dtype).float()) if operand_dtype is not None else (lambda t: t)
    qg, tqg = r_(q.reshape(G, S, D)), r_(tq.reshape(G, S, D))
    kT, tkT = r_(k.reshape(G, Sk, D)).transpose(1, 2), r_(tk.reshape(G, Sk, D)).transpose(1, 2)
    vg, tvg = r_(v.reshape(G, Sk, D)), r_(tv.reshape(G, Sk, D))
    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E

    m = torch.full((G, S, 1), float("-inf"))          # :130
    l = torch.zeros((G, S, 1))                        # :131 (init 0 here, unlike the other paths)
    O = torch.zeros((G, S, D))
    r = torch.zeros((G, S, 1))
    A = torch.zeros((G, S, D))
    Bm = torch.zeros((G, S, D))
    for kb in range(0, Sk, tile_k):
        ke = min(kb + tile_k, Sk)
        Sx = torch.bmm(qg, kT[:, :, kb:ke])                                   # :148
        tS = torch.bmm(tqg, kT[:, :, kb:ke]) + torch.bmm(qg, tkT[:, :, kb:ke])  # :149-152
        tS = tS * sm_scale                                                    # :153 (natural scale)
        m_new = torch.max(m, torch.amax(Sx, -1, keepdim=True) * qk_scale)     # :155-158
        P = torch.exp2(Sx * qk_scale - m_new)                                 # :160-161
        resc = torch.exp2(m - m_new)                                          # :164
        l = l * resc + torch.sum(P, -1, keepdim=True)                         # :165
        m = m_new
        O = O * resc
        O = O + torch.bmm(r_(P), vg[:, kb:ke])                                # :171
        A = A * resc
        A = A + torch.bmm(r_(P), tvg[:, kb:ke])                               # :173-174
        Hm = P * tS                                                           # :176
        r = r * resc + torch.sum(Hm, dim=-1, keepdim=True)                    # :178
        Bm = Bm * resc
        Bm = Bm + torch.bmm(r_(Hm), vg[:, kb:ke])                             # :180-181
    lse = m.squeeze(-1) + torch.log2(l).squeeze(-1)                           # :183
    Of = O / l                                                                # :188
    tO = (A + Bm - r * Of) / l                                                # :190
    return Of.view(B, H, S, D), tO.view(B, H, S, D), lse
