"""Eager definition of the NVFP4 (microscaling) SageAttention3-style forward.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference names FP4 microscaling as the SageAttention3 feature it does NOT
ship (README.md:48-54), so there is no reference code to pin against: parity unpinned.  This file states the contract the
CUDA kernels (csrc/quant_fp4.cu, csrc/attn_fp4_fwd.cu) implement:

  two-level scale   sg = amax_head / 2688 (fp32),  sf = e4m3_rn(amax_blk16 / 6 / sg),  x4 = e2m1_rn(x / (sf * sg))
                    Q, K (K minus its token mean first, one fp16 rounding): blocks of 16 along D;  V: blocks of 16 KEYS
  logits (log2)     u = float(Q4 K4^T with block scales) * (sgq * sgk * sm_scale * log2 e)           per step of 64 / 128 keys
  online softmax    m' = max(m, rowmax u),  P = exp2(u - m'),  l = l * 2^(m - m') + sum P              fp32
  P microscaling    sfp = e4m3_rn(amax_blk16(P) * 448),  P4 = e2m1_rn(P * 2688 / sfp)                  per row, 16 keys
  O                 O = O * 2^(m - m') + (P4 sfp)(V4 sfv);   out = O * sgv / (2688 * l),  lse = m' + log2 l
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504
E2M1_GRID = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])


def e2m1_rn(y: torch.Tensor):
    """Round fp32 to the nearest e2m1 value (ties to the even code, saturating at 6).  Returns (values fp32, codes uint8)."""
    a = y.abs().float()
    # midpoints between neighbours; on a tie the even code (0, 1.0, 2.0, 4.0) wins
    code = ((a > 0.25).to(torch.uint8) + (a >= 0.75).to(torch.uint8) + (a > 1.25).to(torch.uint8) + (a >= 1.75).to(torch.uint8)
            + (a > 2.5).to(torch.uint8) + (a >= 3.5).to(torch.uint8) + (a > 5.0).to(torch.uint8))
    val = E2M1_GRID[code.long()]
    sign = torch.signbit(y)
    return torch.where(sign, -val, val), code | (sign.to(torch.uint8) << 3)


def e4m3_rn(x: torch.Tensor):
    """fp32 (0 <= x <= 448) -> (e4m3 values as fp32, raw bytes uint8)."""
    f8 = x.float().clamp(0.0, 448.0).to(torch.float8_e4m3fn)
    return f8.float(), f8.view(torch.uint8)


def quant_nvfp4(x: torch.Tensor):
    """x: fp32 [G, R, C] (blocks of 16 along the last axis, one two-level scale per G) ->
    (dequantised fp32 [G,R,C], codes uint8 [G,R,C], sf bytes uint8 [G,R,C/16], sg fp32 [G])."""
    G, R, C = x.shape
    sg = (x.abs().amax(dim=(1, 2)).float() / 2688.0)
    xb = x.float().reshape(G, R, C // 16, 16)
    am = xb.abs().amax(dim=-1)
    safe = sg.view(G, 1, 1) > 0
    sf_in = torch.where(safe, (am / 6.0) / sg.view(G, 1, 1), torch.zeros_like(am))
    sf_v, sf_b = e4m3_rn(sf_in)
    scale = sf_v * sg.view(G, 1, 1)                                                # fp32 multiply
    y = torch.where(scale[..., None] > 0, xb / scale[..., None], torch.zeros_like(xb))
    qv, qc = e2m1_rn(y)
    deq = (qv * scale[..., None]).reshape(G, R, C)
    return deq, qc.reshape(G, R, C), sf_b, sg


def pack_codes(codes: torch.Tensor) -> torch.Tensor:
    """uint8 codes [..., C] -> packed bytes [..., C/2], element 2i in the low nibble."""
    return codes[..., 0::2] | (codes[..., 1::2] << 4)


def sf_atoms(sf: torch.Tensor) -> torch.Tensor:
    """sf bytes [T*128, NBLK] (row-major over 128-row tiles) -> the tcgen05.cp atom layout [T, NBLK/4, 512]:
    byte 16 * (r % 32) + 4 * (r / 32) + s of atom (t, k) = scale of row r of tile t, block 4k + s."""
    n, nb = sf.shape
    t = sf.reshape(n // 128, 4, 32, nb // 4, 4)                                    # [tile, r / 32, r % 32, k step, s]
    return t.permute(0, 3, 2, 1, 4).reshape(n // 128, nb // 4, 512).contiguous()


def quantise_inputs(q, k, v):
    """fp16 [B,H,S,D] -> dict of the quantised operands exactly as the CUDA pre-passes emit them."""
    B, H, S_valid, D = q.shape
    G = B * H
    k_mean = k.float().mean(dim=2, keepdim=True).to(torch.float16)                 # LEDGER I-1 (over the valid keys)
    ks = (k - k_mean)                                                              # one fp16 rounding
    S = (S_valid + 127) // 128 * 128                                               # ragged: zero rows up to the next tile boundary
    if S != S_valid:                                                               # (zeros change no amax and no block scale)
        pad = lambda t: torch.cat([t, t.new_zeros((B, H, S - S_valid, D))], dim=2)
        q, ks, v = pad(q), pad(ks), pad(v)
    qd, qc, qsf, sgq = quant_nvfp4(q.reshape(G, S, D).float())
    kd, kc, ksf, sgk = quant_nvfp4(ks.reshape(G, S, D).float())
    vd_t, vc_t, vsf, sgv = quant_nvfp4(v.reshape(G, S, D).float().transpose(1, 2).contiguous())   # [G, D, S]: blocks along keys
    # V scale factors: per 128-key tile one [D rows, 8 blocks] group -> atoms [G * S/128, 2, 512]
    vsf_t = vsf.reshape(G, D, S // 128, 8).permute(0, 2, 1, 3).reshape(G * (S // 128) * D, 8)
    return {"k_mean": k_mean, "qd": qd, "kd": kd, "vd": vd_t.transpose(1, 2).contiguous(),
            "q4": pack_codes(qc).reshape(G * S, D // 2), "k4": pack_codes(kc).reshape(G * S, D // 2), "vt4": pack_codes(vc_t),
            "sfq": sf_atoms(qsf.reshape(G * S, D // 16)), "sfk": sf_atoms(ksf.reshape(G * S, D // 16)), "sfv": sf_atoms(vsf_t),
            "sgq": sgq, "sgk": sgk, "sgv": sgv}


def fp4_fwd(q, k, v, step: int = 128, causal: bool = False):
    """q, k, v fp16 [B,H,S,D] -> (O fp16 [B,H,S,D], lse fp32 [B*H, S] (log2 domain), quantised operands).
    step: keys per online-softmax step (128: the default kernel; 64: the two-CTA variant).  causal: the STRICT mask of the
    reference's baseline (key < query, attention_int8.py:465-473); row 0 of a head = average of the de-quantised V over all keys,
    lse = -128 + log2 S (the int8 path's convention)."""
    B, H, S_valid, D = q.shape
    G = B * H
    S = (S_valid + 127) // 128 * 128
    sm_scale = 1.0 / math.sqrt(D)
    if D == 64:                                                                    # runs with 64 zero columns (block scales unaffected)
        q, k, v = [torch.nn.functional.pad(t, (0, 64)) for t in (q, k, v)]
    qi = quantise_inputs(q, k, v)
    qd, kd, vd = qi["qd"] / qi["sgq"].view(G, 1, 1).clamp_min(1e-38), qi["kd"] / qi["sgk"].view(G, 1, 1).clamp_min(1e-38), \
        qi["vd"] / qi["sgv"].view(G, 1, 1).clamp_min(1e-38)                       # the tensor core sees code * sf only
    c = (qi["sgq"] * qi["sgk"]).view(G, 1, 1) * (sm_scale * LOG2E)
    O = torch.zeros((G, S, qi["vd"].shape[-1]))
    l = torch.zeros((G, S, 1))
    m = torch.full((G, S, 1), float("-inf"))
    for j in range(S // step):
        ks = slice(j * step, (j + 1) * step)
        u = torch.matmul(qd, kd[:, ks].transpose(1, 2)) * c
        keep = (torch.arange(j * step, (j + 1) * step) < S_valid)[None, None, :].expand(1, S, step)     # ragged: padding keys
        if causal:
            keep = keep & (torch.arange(j * step, (j + 1) * step)[None, :] < torch.arange(S)[:, None])[None]
        if causal or S != S_valid:
            u = torch.where(keep, u, torch.full_like(u, float("-inf")))
        m_new = torch.max(m, u.amax(-1, keepdim=True))
        P = torch.exp2(u - m_new)
        resc = torch.exp2(m - m_new)
        if causal or S != S_valid:
            P = torch.where(keep, P, torch.zeros_like(P))                      # (row 0: (-inf) - (-inf))
            resc = torch.where(torch.isinf(m_new), torch.ones_like(resc), resc)
        l = l * resc + P.sum(-1, keepdim=True)
        Pb = P.reshape(G, S, step // 16, 16)
        sfp_v, _ = e4m3_rn(Pb.amax(-1) * 448.0)
        y = torch.where(sfp_v[..., None] > 0, Pb * 2688.0 / sfp_v[..., None], torch.zeros_like(Pb))
        pq, _ = e2m1_rn(y)
        Pd = (pq * sfp_v[..., None]).reshape(G, S, step)
        O = O * resc + torch.matmul(Pd, vd[:, ks])
        m = m_new
    out = O * (qi["sgv"].view(G, 1, 1) / 2688.0) / l
    lse = (m + torch.log2(l)).reshape(G, S)
    if causal:
        out[:, 0] = qi["vd"].sum(dim=1) / S_valid
        lse[:, 0] = -128.0 + math.log2(S_valid)
    return out[:, :S_valid, :D].to(torch.float16).reshape(B, H, S_valid, D), lse[:, :S_valid], qi
