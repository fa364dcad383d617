"""Eager restatement of the int8 SageAttention3-style path (attention_int8.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Pinned bit-for-bit against the unmodified
reference (run under oracle/_helion_standin) by tests/test_oracle_golden.py.

Modes:
  literal  -- what the source does, including attending across the flattened B*H*S axis (I-2)
              and the broken backward (I-5..I-10).
  contract -- per-(b,h) attention; K-smoothing with the per-head token mean (I-1); backward
              with dS = P*(dP - delta), sm_scale, accumulation over tiles, fp32 delta pre-pass,
              fp32 lse.  This is what the CUDA kernels implement.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504  # the reference's literal constant (attention_int8.py:153)


# --------------------------------------------------------------------------------------
# block quantisation -- attention_int8.py:178-186 (Q), :188-195 (K), :241-247 (V)
# --------------------------------------------------------------------------------------
def _to_i8(x: torch.Tensor, rounding: str = "trunc") -> torch.Tensor:
    """"trunc": the reference's `.to(torch.int8)` (toward zero, LEDGER I-3); "nearest": round half to even - the opt-in
    accuracy mode of the CUDA path, which removes the truncation bias."""
    assert rounding in ("trunc", "nearest")
    return (torch.round(x.float()) if rounding == "nearest" else x).to(torch.int8)


def quant_block(x2d: torch.Tensor, blk: int, rounding: str = "trunc"):
    """x2d: [N, D] fp16 -> (int8 [N, D], fp16 scales [ceil(N/blk)]).

    scale = amax(|block|)/127 in fp16; value = trunc_toward_zero(fp16(x / scale)).
    All-zero block (reference: 0/0 -> NaN -> undefined int8, LEDGER I-4) is DEFINED as
    scale 0, values 0.
    """
    assert x2d.dtype == torch.float16 and x2d.dim() == 2
    n, d = x2d.shape
    nblk = (n + blk - 1) // blk
    out = torch.empty((n, d), dtype=torch.int8)
    scales = torch.empty((nblk,), dtype=torch.float16)
    if n % blk == 0:
        xb = x2d.reshape(nblk, blk * d)
        s = torch.amax(xb.abs(), dim=1) / 127                     # fp16 (:180)
        qv = xb / s[:, None]                                      # fp16 divide (:182)
        qv = torch.where(s[:, None] == 0, torch.zeros_like(qv), qv)
        out.copy_(_to_i8(qv, rounding).reshape(n, d))             # trunc (:183)
        scales.copy_(s)
        return out, scales
    for b in range(nblk):                                         # ragged tail (hl.tile clamps)
        blkx = x2d[b * blk:(b + 1) * blk]
        s = torch.amax(blkx.abs().flatten(), dim=0) / 127
        qv = blkx / s if float(s) != 0.0 else torch.zeros_like(blkx)
        out[b * blk:(b + 1) * blk] = _to_i8(qv, rounding)
        scales[b] = s
    return out, scales


def k_token_mean(k: torch.Tensor) -> torch.Tensor:
    """Contract K-smoothing mean (LEDGER I-1): per-(b,h) mean over tokens, fp32 accumulate,
    rounded to fp16, shape [B,H,1,D].  (The literal `k.mean(0)[:, :, :, None]`,
    attention_int8.py:24-25, raises on 4-D input.)"""
    return k.float().mean(dim=2, keepdim=True).to(torch.float16)


def smooth_k(k: torch.Tensor, k_mean: torch.Tensor) -> torch.Tensor:
    """k - mean in fp16 (one rounding), attention_int8.py:25."""
    return k - k_mean


def _imm(a_i8: torch.Tensor, b_i8: torch.Tensor) -> torch.Tensor:
    """Exact int8 x int8 -> int32 matmul through fp32 (|sum| <= 127*127*K < 2^24 for K<=1024)."""
    assert a_i8.shape[-1] <= 1024
    return torch.matmul(a_i8.to(torch.float32), b_i8.to(torch.float32)).to(torch.int32)


# --------------------------------------------------------------------------------------
# forward -- attention_int8.py:170-257
# --------------------------------------------------------------------------------------
def int8_fwd(q, k_smooth, v, Bq: int = 32, Bkv: int = 32, per_head: bool = True,
             return_lse32: bool = False, rounding: str = "trunc", causal: bool = False, s_valid: int | None = None):
    """Returns the reference 10-tuple
        (O fp16 [B,H,S,D], lse fp16 [N], q_i8 [N,D], k_i8_T [D,N], v_i8 [N,D],
         sq [N/Bq], sk [N/Bkv], sv [N/Bkv], Bq, Bkv)
    per_head=False reproduces the literal flattened attention (LEDGER I-2).
    Every dtype/rounding step follows the table in SURVEY.md 3.1.
    causal=True (absent in the reference; contract of the CUDA path): the strict mask of the
    reference's own baseline (key < query, attention_int8.py:465-473) with weight exactly 0; row maxima, P scales and
    row sums run over the visible keys only; a (row, tile) pair without a visible key contributes nothing; row 0 of a
    head, which sees no key at all, is the uniform average over ALL keys of the de-quantised V with
    lse = -128 + log2(S) (LEDGER B-1, what the baseline's finite fill value produces).
    s_valid (ragged sequences; the reference's hl.tile clamps the last tile, attention_int8.py:170,176): q, k, v are
    zero-padded per head to a multiple of the block sizes, keys >= s_valid have weight exactly 0 and k-tiles without a
    valid key are skipped; rows >= s_valid of the outputs are padding.
    """
    B, H, S, D = q.shape
    N = B * H * S
    q_i8, sq = quant_block(q.reshape(N, D), Bq, rounding)
    k_i8, sk = quant_block(k_smooth.reshape(N, D), Bkv, rounding)
    v_i8, sv = quant_block(v.reshape(N, D), Bkv, rounding)

    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E

    G, L = (B * H, S) if per_head else (1, N)      # groups that attend independently
    assert L % Bq == 0 and L % Bkv == 0, "oracle fast path needs S % Bq == S % Bkv == 0"
    qg = q_i8.view(G, L, D)
    kg = k_i8.view(G, L, D)
    vg = v_i8.view(G, L, D)
    sq_rows = sq.view(G, L // Bq).repeat_interleave(Bq, dim=1)[..., None].float()   # [G,L,1]
    sk_g = sk.view(G, L // Bkv)
    sv_g = sv.view(G, L // Bkv)

    O = torch.zeros((G, L, D), dtype=torch.float32)
    l = torch.full((G, L, 1), 1.0, dtype=torch.float32)                 # :173 (init 1.0)
    m = torch.full((G, L, 1), float("-inf"), dtype=torch.float16)       # :174

    for j in range(L // Bkv if s_valid is None else -(-s_valid // Bkv)):
        ks = slice(j * Bkv, (j + 1) * Bkv)
        acc = _imm(qg, kg[:, ks].transpose(1, 2))                        # :197
        skj = sk_g[:, j].view(G, 1, 1).float()
        S32 = acc.to(torch.float32) * sq_rows * skj * qk_scale           # :200
        S16 = S32.to(torch.float16)                                      # :203
        if s_valid is not None:
            pad = torch.arange(j * Bkv, (j + 1) * Bkv) >= s_valid
            S16 = torch.where(pad[None, None, :], torch.tensor(float("-inf"), dtype=torch.float16), S16)
        if causal:
            assert per_head
            keep = torch.arange(L)[:, None] > torch.arange(j * Bkv, (j + 1) * Bkv)[None, :]     # strict: key < query
            S16 = torch.where(keep, S16, torch.tensor(float("-inf"), dtype=torch.float16))
        row_max = torch.amax(S16, -1, keepdim=True)                      # :205
        m_new = torch.max(m, row_max)                                    # :206-209
        P = torch.exp2((S16 - m_new).to(torch.float32))                  # :211-213 (fp16 subtract)
        rescale = torch.exp2((m - m_new).to(torch.float32))              # :217-219
        sp = torch.exp2((row_max - m_new).to(torch.float32)) / 127       # :232-234
        if causal:                                                       # (-inf) - (-inf): nothing visible yet / in this tile
            P = torch.where(keep, P, torch.zeros_like(P))
            rescale = torch.where(torch.isinf(m_new), torch.ones_like(rescale), rescale)
        l_new = torch.sum(P, -1, keepdim=True)                           # :215
        m = m_new
        l = l * rescale + l_new                                          # :223
        O = O * rescale                                                  # :225
        Pq = P / sp
        if causal:
            Pq = torch.where(torch.isinf(row_max), torch.zeros_like(Pq), Pq)
        P_i8 = _to_i8(Pq, rounding)                                      # :236-237
        svj = sv_g[:, j].view(G, 1, 1).float()
        O = O + _imm(P_i8, vg[:, ks]).to(torch.float32) * sp * svj       # :249-250

    lse32 = m.squeeze(-1).float() + torch.log2(l).squeeze(-1)
    lse16 = m.squeeze(-1) + torch.log2(l).squeeze(-1).to(torch.float16)  # :252
    O16 = (O / l).to(torch.float16)                                      # :256-257
    if causal:                                                           # row 0 of every head (LEDGER B-1)
        v_deq = vg.float() * sv_g.repeat_interleave(Bkv, dim=1)[..., None].float()
        O16[:, 0] = v_deq.mean(dim=1).to(torch.float16)
        lse32[:, 0] = -128.0 + math.log2(L)
        lse16[:, 0] = torch.tensor(-128.0 + math.log2(L), dtype=torch.float16)
    out = (O16.view(B, H, S, D), lse16.reshape(N), q_i8, k_i8.t(), v_i8, sq, sk, sv, Bq, Bkv)
    if return_lse32:
        return out + (lse32.reshape(N),)
    return out


def int8_attend_state(q_i8, sq, k_i8, v_i8, sk, sv, state, BH, Sq, Sk, D, Bq, Bkv, last, causal_diag: bool = False):
    """One ring step on pre-quantised operands: the tile loop of `int8_fwd` (attention_int8.py:176-250) over one K/V
    shard, continuing the online-softmax state (O_acc fp32 [BH*Sq,D], m fp32 (holding an fp16 value), l fp32).
    last=False -> new state; last=True -> (O fp16 [BH*Sq,D], lse16, lse32) as :252-257.
    causal_diag=True (the diagonal chunk of a causal ring, Sq == Sk): strict mask key < query inside the chunk with weight
    exactly 0, as `int8_fwd(causal=True)`; a row without a visible key keeps its state (row 0 of a fresh chunk: O = 0,
    m = -inf, l = 1) - the row-0 rule of LEDGER B-1 is applied by the caller."""
    qk_scale = (1.0 / math.sqrt(D)) * LOG2E
    qg, kg, vg = q_i8.view(BH, Sq, D), k_i8.view(BH, Sk, D), v_i8.view(BH, Sk, D)
    sq_rows = sq.view(BH, Sq // Bq).repeat_interleave(Bq, dim=1)[..., None].float()
    sk_g, sv_g = sk.view(BH, Sk // Bkv), sv.view(BH, Sk // Bkv)
    if state is None:
        O = torch.zeros((BH, Sq, D), dtype=torch.float32)
        l = torch.full((BH, Sq, 1), 1.0, dtype=torch.float32)
        m = torch.full((BH, Sq, 1), float("-inf"), dtype=torch.float16)
    else:
        O = state[0].view(BH, Sq, D).clone()
        m = state[1].view(BH, Sq, 1).to(torch.float16)
        l = state[2].view(BH, Sq, 1).clone()
    for j in range(Sk // Bkv):
        ks = slice(j * Bkv, (j + 1) * Bkv)
        S16 = (_imm(qg, kg[:, ks].transpose(1, 2)).to(torch.float32) * sq_rows * sk_g[:, j].view(BH, 1, 1).float()
               * qk_scale).to(torch.float16)
        if causal_diag:
            keep = torch.arange(Sq)[:, None] > torch.arange(j * Bkv, (j + 1) * Bkv)[None, :]
            if not bool(keep.any()):
                continue                                                # tile above the diagonal: skipped
            S16 = torch.where(keep, S16, torch.tensor(float("-inf"), dtype=torch.float16))
        row_max = torch.amax(S16, -1, keepdim=True)
        m_new = torch.max(m, row_max)
        P = torch.exp2((S16 - m_new).to(torch.float32))
        rescale = torch.exp2((m - m_new).to(torch.float32))
        if causal_diag:                                                 # (-inf) - (-inf): nothing visible yet / in this tile
            P = torch.where(keep, P, torch.zeros_like(P))
            rescale = torch.where(torch.isinf(m_new), torch.ones_like(rescale), rescale)
        m = m_new
        l = l * rescale + torch.sum(P, -1, keepdim=True)
        O = O * rescale
        sp = torch.exp2((row_max - m).to(torch.float32)) / 127
        Pq = P / sp
        if causal_diag:
            Pq = torch.where(torch.isinf(row_max), torch.zeros_like(Pq), Pq)
            sp = torch.where(torch.isinf(row_max), torch.zeros_like(sp), sp)
        P_i8 = Pq.to(torch.int8)
        O = O + _imm(P_i8, vg[:, ks]).to(torch.float32) * sp * sv_g[:, j].view(BH, 1, 1).float()
    if not last:
        return O.reshape(BH * Sq, D), m.float().reshape(-1), l.reshape(-1)
    lse32 = m.squeeze(-1).float() + torch.log2(l).squeeze(-1)
    lse16 = m.squeeze(-1) + torch.log2(l).squeeze(-1).to(torch.float16)
    return (O / l).to(torch.float16).reshape(BH * Sq, D), lse16.reshape(-1), lse32.reshape(-1)


def sage_forward(q, k, v, Bq=32, Bkv=32, rounding: str = "trunc", causal: bool = False):
    """Contract version of SageAttention3_Int8_autograd_function.forward
    (attention_int8.py:21-40 with LEDGER I-1): 11-tuple with k_mean [B,H,1,D] in slot 2."""
    km = k_token_mean(k)
    out = int8_fwd(q, smooth_k(k, km), v, Bq, Bkv, per_head=True, rounding=rounding, causal=causal)
    return out[:2] + (km,) + out[2:]


# --------------------------------------------------------------------------------------
# backward -- attention_int8.py:342-428
# --------------------------------------------------------------------------------------
def int8_bwd_literal(dO, q_i8, sq, k_i8_T, k_mean_bhk, sk, v_i8, sv, O, lse16, Bq, Bkv):
    """Bug-for-bug restatement (flattened tiles, dS = S*(dP-D), overwrites, qk_scale, per-token
    k_mean scalar).  `k_mean_bhk` is the 3-D [batch, head, tokens] tensor the literal code
    expects (:304).  Tiny shapes only (python tile loops)."""
    N, D = q_i8.shape
    B, H, S, _ = O.shape
    O_bh = O.reshape(N, D)
    dO_bh = dO.reshape(N, D)
    kmean = k_mean_bhk.reshape(-1)
    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E
    dq = torch.zeros((N, D), dtype=torch.float16)
    dk = torch.zeros((N, D), dtype=torch.float16)
    dv = torch.zeros((N, D), dtype=torch.float16)
    for jt in range((N + Bkv - 1) // Bkv):
        ks = slice(jt * Bkv, min((jt + 1) * Bkv, N))
        for it in range((N + Bq - 1) // Bq):
            qs = slice(it * Bq, min((it + 1) * Bq, N))
            k_T = k_i8_T[:, ks]
            q_b = q_i8[qs]
            acc = _imm(q_b, k_T)                                                   # :352
            S16 = (acc.to(torch.float32) * sq[it] * sk[jt] * qk_scale).to(torch.float16)  # :353-355
            l = lse16[qs]
            P = torch.exp2((S16 - l[:, None]).to(torch.float32))                   # :360
            sP = torch.amax(P.abs().flatten(), dim=0) / 127                        # :363
            P_i8 = (P / sP).to(torch.int8)
            dO_b = dO_bh[qs]
            s_dO = torch.amax(dO_b.abs().flatten(), dim=0) / 127                   # :372 (fp16)
            dO_i8 = (dO_b / s_dO).to(torch.int8)
            dv_t = (_imm(P_i8.t(), dO_i8).to(torch.float32) * s_dO * sP).to(torch.float16)  # :375-378
            dP = _imm(dO_i8, v_i8[ks].t()).to(torch.float32) * s_dO * sv[jt]       # :382-384
            Dv = torch.sum(dO_b * O_bh[qs], dim=-1, keepdim=True)                  # :398 (fp16)
            dS = S16.to(torch.float32) * (dP - Dv.to(torch.float32))               # :399
            s_dS = torch.amax(dS.abs().flatten(), dim=0) / 127                     # :403
            dS_i8 = (dS / s_dS).to(torch.int8)
            dSk = torch.sum(dS, dim=-1) * kmean[qs]                                # :409
            dSK = _imm(dS_i8, k_T.t()).to(torch.float32) * s_dS * sk[jt] * qk_scale  # :416-417
            dq[qs] += dSK + dSk.to(torch.float32)[:, None]                         # :420
            dk_t = (_imm(dS_i8.t(), q_b).to(torch.float32) * s_dS * sq[it] * qk_scale).to(torch.float16)
            dk[ks] = dk_t                                                          # :427
            dv[qs] = dv_t                                                          # :428
    return dq.view(B, H, S, D), dk.view(B, H, S, D), dv.view(B, H, S, D)


def quant_tile_fp32(x: torch.Tensor, rounding: str = "trunc"):
    """Per-[Bq,Bkv]-tile quantisation of an fp32 tile batch x: [..., r, c] (attention_int8.py
    :363-365, :403-405): scale = amax|x|/127 (fp32), value = trunc(x/scale); zero tile -> 0."""
    s = torch.amax(x.abs(), dim=(-2, -1), keepdim=True) / 127
    qv = torch.where(s == 0, torch.zeros_like(x), x / s)
    return _to_i8(qv, rounding), s


def int8_bwd_contract(dO, q_i8, sq, k_i8_T, k_mean, sk, v_i8, sv, O, lse, Bq, Bkv, rounding: str = "trunc",
                      causal: bool = False, s_valid: int | None = None, v_fp16=None):
    """CONTRACT backward (what the CUDA kernel implements; LEDGER I-1,5,6,7,8,9,10,12,15).

    Per (b,h); k-tile j, q-tile i:
      S16 = fp16(float(q_i8 k_i8^T) * sq * sk * qk_scale)            (as forward)
      P   = exp2(float(S16) - lse32)                                  (lse may be fp16 or fp32)
      P_i8, sP   = per-tile quant;  dO_i8, s_dO = per-[Bq,D]-block quant (pre-pass, fp16 rule)
      dV[j] += float(P_i8^T dO_i8) * s_dO * sP
      dP   = float(dO_i8 v_i8^T) * s_dO * sv
      dS   = P * (dP - delta),  delta = rowsum(float(dO)*float(O)) fp32 pre-pass
      dS_i8, s_dS = per-tile quant
      dQ[i] += float(dS_i8 k_i8) * s_dS * sk * sm_scale + sm_scale * rowsum(dS) x k_mean
      dK[j] += float(dS_i8^T q_i8) * s_dS * sq * sm_scale
    fp32 accumulation, fp16 outputs.  k_mean: [B,H,1,D] fp16.
    causal=True: tiles with q-tile < k-tile are skipped, masked P is exactly 0 (so dS is 0 there and the
    tile-wide amax runs over the visible entries); row 0 of a head attends uniformly to all S keys (LEDGER B-1), which
    adds dO[0]/S to every dV row and nothing to dQ / dK.
    """
    N, D = q_i8.shape
    B, H, S, _ = O.shape
    G = B * H
    assert S % Bq == 0 and S % Bkv == 0
    sm_scale = 1.0 / math.sqrt(D)
    qk_scale = sm_scale * LOG2E
    nq, nk = S // Bq, S // Bkv
    qg = q_i8.view(G, nq, Bq, D)
    kg = k_i8_T.t().reshape(G, nk, Bkv, D)
    vg = v_i8.view(G, nk, Bkv, D)
    sqg = sq.view(G, nq).float()
    skg = sk.view(G, nk).float()
    svg = sv.view(G, nk).float()
    lse32 = lse.float().view(G, nq, Bq, 1)
    dOg = dO.reshape(G, nq, Bq, D)
    delta = (dOg.float() * O.reshape(G, nq, Bq, D).float()).sum(-1, keepdim=True)   # fp32 pre-pass
    dO_i8, s_dO = quant_block(dO.reshape(N, D).to(torch.float16), Bq, rounding)
    dO_i8 = dO_i8.view(G, nq, Bq, D)
    s_dO = s_dO.view(G, nq).float()
    km = k_mean.reshape(G, 1, D).float()

    dq = torch.zeros((G, nq, Bq, D), dtype=torch.float32)
    dk = torch.zeros((G, nk, Bkv, D), dtype=torch.float32)
    dv = torch.zeros((G, nk, Bkv, D), dtype=torch.float32)
    for j in range(nk):
        kj = kg[:, j]                                              # [G,Bkv,D]
        vj = vg[:, j]
        for i in range(nq):
            if causal and i * Bq + Bq - 1 <= j * Bkv:                # no (query, key) pair with key < query
                continue
            qi = qg[:, i]
            acc = _imm(qi, kj.transpose(1, 2))
            S16 = (acc.to(torch.float32) * sqg[:, i, None, None] * skg[:, j, None, None] * qk_scale).to(torch.float16)
            P = torch.exp2(S16.to(torch.float32) - lse32[:, i])
            if s_valid is not None:                                # padded keys of a ragged sequence: P = 0
                P = torch.where((torch.arange(j * Bkv, (j + 1) * Bkv) >= s_valid)[None, None, :], torch.zeros_like(P), P)
            if causal:
                keep = torch.arange(i * Bq, (i + 1) * Bq)[:, None] > torch.arange(j * Bkv, (j + 1) * Bkv)[None, :]
                P = torch.where(keep, P, torch.zeros_like(P))
            P_i8, sP = quant_tile_fp32(P, rounding)
            dv[:, j] += _imm(P_i8.transpose(1, 2), dO_i8[:, i]).to(torch.float32) * s_dO[:, i, None, None] * sP
            if v_fp16 is None:
                dP = _imm(dO_i8[:, i], vj.transpose(1, 2)).to(torch.float32) * s_dO[:, i, None, None] * svg[:, j, None, None]
            else:                                                  # SageBwd option: dO V^T from the fp16 tensors
                dP = torch.matmul(dOg[:, i].float(), v_fp16.reshape(G, nk, Bkv, D)[:, j].float().transpose(1, 2))
            dS = P * (dP - delta[:, i])
            dS_i8, s_dS = quant_tile_fp32(dS, rounding)
            dq[:, i] += _imm(dS_i8, kj).to(torch.float32) * s_dS * skg[:, j, None, None] * sm_scale \
                + sm_scale * dS.sum(-1, keepdim=True) * km
            dk[:, j] += _imm(dS_i8.transpose(1, 2), qi).to(torch.float32) * s_dS * sqg[:, i, None, None] * sm_scale
    if causal:                                                     # row 0: uniform over all keys
        dv = dv + (dO.reshape(G, S, D)[:, :1].float() / S).view(G, 1, 1, D)
    f = lambda t: t.to(torch.float16).view(B, H, S, D)
    return f(dq), f(dk), f(dv)
