"""Opt-in accuracy mode: round-to-nearest instead of the reference's truncation in every int8 quantiser
of the path.  Same kernels, one instruction modifier; checked against the oracle's `rounding="nearest"` switch and
against fp32 math (the gradient error must drop)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("blk,D", [(128, 128), (32, 64), (256, 128)])
def test_quant_block_nearest_bit_exact(blk, D):
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    g = torch.Generator().manual_seed(11 + blk)
    x = (torch.randn(4 * 512, D, generator=g) * 3).half()
    x[:blk] = 0                                                    # all-zero block (LEDGER I-4)
    qi, s = ops.quant_block(x.cuda(), blk, rounding="nearest")
    qr, sr = int8_ref.quant_block(x, blk, rounding="nearest")
    assert torch.equal(qi.cpu(), qr) and torch.equal(s.cpu(), sr)
    qt, _ = ops.quant_block(x.cuda(), blk)                         # default stays the reference's truncation
    assert torch.equal(qt.cpu(), int8_ref.quant_block(x, blk)[0]) and not torch.equal(qt.cpu(), qr)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_nearest_mode_forward_backward_vs_oracle_and_fp32():
    from oracle import int8_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = 1, 2, 512, 128
    g = torch.Generator().manual_seed(99)
    q, k, v, dO = [torch.randn(B, H, S, D, generator=g).half() for _ in range(4)]
    # fp32 ground truth
    qf, kf, vf = [t.float().requires_grad_() for t in (q, k, v)]
    Of = baseline_pytorch_attention(qf, kf, vf, D, False)
    Of.backward(dO.float())
    res = {}
    try:
        for mode in ("trunc", "nearest"):
            A.set_quant_rounding(mode)
            qr, kr, vr = [t.cuda().requires_grad_() for t in (q, k, v)]
            O = A.sage_attention_3_int8(qr, kr, vr)
            O.backward(dO.cuda())
            torch.cuda.synchronize()
            res[mode] = (O.detach().cpu(), qr.grad.cpu(), kr.grad.cpu(), vr.grad.cpu())
            # same-mode oracle (contract): forward tuple + backward
            out = int8_ref.sage_forward(q, k, v, 128, 128, rounding=mode)
            fwd = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, out[2]), v, 128, 128, return_lse32=True, rounding=mode)
            assert (res[mode][0].float() - out[0].float()).abs().max() < 5e-3
            dq, dk, dv = int8_ref.int8_bwd_contract(dO, out[3], out[6], out[4], out[2], out[7], out[5], out[8], out[0],
                                                    fwd[-1], 128, 128, rounding=mode)
            for got, ref in zip(res[mode][1:], (dq, dk, dv)):
                assert _rel(got, ref) < 3e-2
    finally:
        A.set_quant_rounding("trunc")
    truth = (Of.detach(), qf.grad, kf.grad, vf.grad)
    err = {m: [_rel(a, b) for a, b in zip(res[m], truth)] for m in res}
    # nearest rounding is unbiased: every output gets closer to fp32 math, the gradients markedly so
    for i in range(4):
        assert err["nearest"][i] < err["trunc"][i], err
    assert sum(err["nearest"][1:]) < 0.75 * sum(err["trunc"][1:]), err


def test_set_quant_rounding_validates():
    from quantizedattention_b200 import attention_int8 as A
    with pytest.raises(ValueError):
        A.set_quant_rounding("up")
