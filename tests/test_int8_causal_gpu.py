"""Causal int8 attention (absent in the reference's int8 kernel): strict mask of the reference's own
baseline (key < query), masked weight exactly 0, fully masked tiles skipped, row 0 of a head = uniform average over all
keys.  Checked against the oracle's `causal=True` contract and against fp32 math with the same mask."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("shape", [(1, 2, 512, 128), (2, 2, 384, 64), (1, 1, 128, 128), (1, 1, 1024, 64)])
def test_int8_causal_forward_backward(shape):
    from oracle import int8_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = shape
    g = torch.Generator().manual_seed(31 + S + D)
    q, k, v, dO = [torch.randn(shape, generator=g).half() for _ in range(4)]
    qr, kr, vr = [t.cuda().requires_grad_() for t in (q, k, v)]
    O = A.sage_attention_3_int8(qr, kr, vr, causal=True)
    O.backward(dO.cuda())
    torch.cuda.synchronize()
    # ---- same-contract oracle
    out = int8_ref.sage_forward(q, k, v, 128, 128, causal=True)
    assert (O.detach().cpu().float() - out[0].float()).abs().max() < 5e-3
    v_deq = out[5].view(B * H, S, D).float() * out[8].view(B * H, S // 128).repeat_interleave(128, 1)[..., None].float()
    assert (O.detach().cpu()[:, :, 0].reshape(B * H, D).float() - v_deq.mean(1)).abs().max() < 2e-3     # row 0 (LEDGER B-1)
    fwd = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, out[2]), v, 128, 128, return_lse32=True, causal=True)
    dq, dk, dv = int8_ref.int8_bwd_contract(dO, out[3], out[6], out[4], out[2], out[7], out[5], out[8], out[0], fwd[-1],
                                            128, 128, causal=True)
    for got, ref, name in zip((qr.grad, kr.grad, vr.grad), (dq, dk, dv), "qkv"):
        assert _rel(got.cpu(), ref) < 3e-2, (name, _rel(got.cpu(), ref))
    assert qr.grad[:, :, 0].abs().max() == 0                       # row 0 sees no key: no gradient to q
    # ---- fp32 math with the same mask (quantisation-level agreement)
    qf, kf, vf = [t.float().requires_grad_() for t in (q, k, v)]
    base = baseline_pytorch_attention(qf, kf, vf, D, True)
    base.backward(dO.float())
    assert (O.detach().cpu().float() - base).abs().max() < 0.15 and _rel(O.detach().cpu(), base.detach()) < 8e-2   # early rows: few keys, coarse P
    for got, ref in zip((qr.grad, kr.grad, vr.grad), (qf.grad, kf.grad, vf.grad)):
        assert _rel(got.cpu(), ref) < 0.2


def test_int8_causal_unsupported_combinations_raise():
    from quantizedattention_b200 import attention_int8 as A
    q = torch.randn(1, 1, 256, 64).half().cuda()
    try:
        A.set_block_sizes(64, 64)
        with pytest.raises(RuntimeError):
            A.sage_attention_3_int8(q, q, q, causal=True)
    finally:
        A.set_block_sizes(128, 128)
    try:
        A.set_quant_rounding("nearest")
        with pytest.raises(RuntimeError):
            A.sage_attention_3_int8(q, q, q, causal=True)
    finally:
        A.set_quant_rounding("trunc")
