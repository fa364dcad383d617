"""Ragged sequence lengths (S not a multiple of 128 / of the block sizes; the reference's hl.tile clamps the last tile,
attention_int8.py:170,176) and variable-length batches for the int8 path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _pad(t, Sp):
    B, H, S, D = t.shape
    out = torch.zeros((B, H, Sp, D), dtype=t.dtype)
    out[:, :, :S] = t
    return out


@pytest.mark.parametrize("shape,Bq,Bkv", [((1, 2, 200, 128), 128, 128), ((2, 2, 321, 64), 128, 128), ((1, 2, 130, 128), 32, 32),
                                          ((1, 1, 100, 64), 64, 64), ((1, 2, 300, 128), 128, 256), ((1, 2, 1000, 128), 128, 128)])
def test_int8_ragged_forward_backward_match_oracle(shape, Bq, Bkv):
    from oracle import int8_ref
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = shape
    g = torch.Generator().manual_seed(6000 + S + D + Bq)
    q, k, v, dO = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(4)]
    k = (k.float() + 0.5).to(torch.float16)
    out = A.SageAttention3_Int8_autograd_function.apply(q.cuda(), k.cuda(), v.cuda(), Bq=Bq, Bkv=Bkv)
    O, lse16, kmean, q_i8, k_i8_T, v_i8, sq, sk, sv, bq, bkv = out
    assert O.shape == shape and q_i8.shape == (B * H * S, D) and k_i8_T.shape == (D, B * H * S)
    assert sq.numel() == B * H * -(-S // Bq) and sk.numel() == B * H * -(-S // Bkv)
    # oracle on zero-padded tensors with the padded keys masked
    Sp = -(-S // max(128, Bq, Bkv)) * max(128, Bq, Bkv)
    km = int8_ref.k_token_mean(k)
    assert torch.equal(kmean.cpu(), km)
    ks = int8_ref.smooth_k(k, km)
    ref = int8_ref.int8_fwd(_pad(q, Sp), _pad(ks, Sp), _pad(v, Sp), Bq, Bkv, per_head=True, return_lse32=True, s_valid=S)
    cut = lambda t2d: t2d.reshape(B * H, Sp, -1)[:, :S].reshape(B * H * S, -1)
    assert torch.equal(q_i8.cpu(), cut(ref[2])) and torch.equal(v_i8.cpu(), cut(ref[4]))
    assert torch.equal(k_i8_T.t().cpu(), cut(ref[3].t()))
    assert torch.equal(sq.cpu(), ref[5].view(B * H, -1)[:, :-(-S // Bq)].reshape(-1))
    Oref = ref[0][:, :, :S]
    assert (O.cpu().float() - Oref.float()).abs().max() < 5e-3 and _cos(O.cpu(), Oref) > 0.99999
    if Bkv == 256:
        return                                                     # the backward runs Bkv in {32, 64, 128}
    got = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv)
    torch.cuda.synchronize()
    # oracle backward on the padded tensors of the oracle forward (dO rows of the padding are zero)
    Op = _pad(O.cpu(), Sp)
    lsep = torch.full((B * H, Sp), 6.0e4, dtype=torch.float16)       # padded query rows: P = exp2(S - lse) = 0
    lsep[:, :S] = lse16.cpu().view(B * H, S)
    rb = int8_ref.int8_bwd_contract(_pad(dO, Sp), ref[2], ref[5], ref[3], km, ref[6], ref[4], ref[7], Op, lsep.reshape(-1), Bq, Bkv,
                                    s_valid=S)
    for name, a, b in zip(("dq", "dk", "dv"), got, rb):
        b = b[:, :, :S]
        assert a.shape == shape and torch.isfinite(a.float()).all(), name
        assert _cos(a.cpu(), b) > 0.9995 and _rel(a.cpu(), b) < 3e-2, (name, _cos(a.cpu(), b), _rel(a.cpu(), b))


def test_int8_ragged_autograd_vs_fp32_math():
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    shape = (1, 3, 459, 128)
    g = torch.Generator().manual_seed(91)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    qh, kh, vh = [t.half().cuda().requires_grad_() for t in (q, k, v)]
    O = A.sage_attention_3_int8(qh, kh, vh)
    O.backward(dO.half().cuda())
    qf, kf, vf = [t.half().float().requires_grad_() for t in (q, k, v)]
    Ob = baseline_pytorch_attention(qf, kf, vf, 128, False)
    Ob.backward(dO.half().float())
    assert (O.detach().cpu().float() - Ob.detach()).abs().max() < 8e-2
    for a, b in zip((qh, kh, vh), (qf, kf, vf)):
        assert _cos(a.grad.cpu(), b.grad) > 0.99 and _rel(a.grad.cpu(), b.grad) < 0.2


def test_int8_varlen_matches_per_sequence_calls():
    from quantizedattention_b200 import attention_int8 as A
    from quantizedattention_b200 import modules as M
    lens = [256, 100, 256, 333]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    H, D = 2, 128
    g = torch.Generator().manual_seed(17)
    q, k, v, dO = [torch.randn(cu[-1], H, D, generator=g).half().cuda() for _ in range(4)]
    qr, kr, vr = [t.clone().requires_grad_() for t in (q, k, v)]
    O = M.sage_attention_int8_varlen(qr, kr, vr, torch.tensor(cu))
    O.backward(dO)
    torch.cuda.synchronize()
    for b, n in enumerate(lens):
        sl = slice(cu[b], cu[b + 1])
        one = [t[sl].transpose(0, 1).unsqueeze(0).contiguous().requires_grad_() for t in (q, k, v)]    # [1,H,S,D]
        Ob = A.sage_attention_3_int8(*one)
        Ob.backward(dO[sl].transpose(0, 1).unsqueeze(0).contiguous())
        assert torch.equal(O[sl], Ob[0].transpose(0, 1))
        assert torch.equal(kr.grad[sl], one[1].grad[0].transpose(0, 1)) and torch.equal(vr.grad[sl], one[2].grad[0].transpose(0, 1))
        assert (qr.grad[sl].float() - one[0].grad[0].transpose(0, 1).float()).abs().max() <= 2e-3 * one[0].grad.abs().max() + 1e-6
    with pytest.raises(ValueError):
        M.sage_attention_int8_varlen(q, k, v, [0, 10, 5, cu[-1]])
