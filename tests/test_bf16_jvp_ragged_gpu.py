"""Sequence lengths that are not a multiple of 128 on the bf16 path (rows a5-a7) and the JVP kernel (row a8): the
reference's hl.tile clamps the last tile (attention_bf16.py:170,201,361; attention_jvp.py:120,137); here the operands are
zero-padded per head and the kernels mask the padded keys.  Compared with the oracle run on the UNPADDED tensors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("S,D", [(200, 128), (333, 64), (1000, 128), (65, 64), (449, 128)])
def test_bf16_fwd_bwd_ragged_matches_oracle(S, D, causal):
    from oracle import bf16_ref
    from quantizedattention_b200 import ops
    shape = (1, 2, S, D)
    g = torch.Generator().manual_seed(6000 + S + D + causal)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    q, k, v = q.half(), k.half(), v.bfloat16()
    O, lse = ops.bf16_fwd(q.cuda(), k.cuda(), v.cuda(), causal)
    torch.cuda.synchronize()
    assert O.shape == shape and lse.shape == (2, S) and torch.isfinite(O).all()
    Sp = (S + 127) // 128 * 128
    Or, lser = bf16_ref.bf16_fwd(q, k, v, causal, tile_k=ops.bf16_fwd_key_step(Sp), mode="contract", lazy_tau=ops.BF16_RESCALE_TAU)
    assert (O.cpu() - Or).abs().max() < 2.5e-2
    assert (lse.cpu() - lser).abs().max() < 2e-2
    for variant in ((0, 1) if D == 128 else (0,)):
        got = ops.bf16_bwd(q.cuda(), k.cuda(), v.cuda(), O, lse, causal, dO.cuda(), variant=variant)
        torch.cuda.synchronize()
        ref = bf16_ref.bf16_bwd(q, k, v, O.cpu(), lse.cpu(), causal, dO, mode="contract")
        for name, a, b in zip(("dq", "dk", "dv"), got, ref):
            assert a.shape == shape and torch.isfinite(a).all(), name
            assert _rel(a.cpu(), b) < 6e-3, (name, variant, _rel(a.cpu(), b))


def test_bf16_cross_attention_ragged_keys():
    """Sq != Sk, both ragged (forward only: the backward is self-attention, LEDGER I-11)."""
    from oracle import bf16_ref
    from quantizedattention_b200 import ops
    g = torch.Generator().manual_seed(61)
    q = torch.randn(1, 2, 300, 128, generator=g).half()
    k = torch.randn(1, 2, 77, 128, generator=g).half()
    v = torch.randn(1, 2, 77, 128, generator=g).bfloat16()
    O, lse = ops.bf16_fwd(q.cuda(), k.cuda(), v.cuda(), False)
    Or, lser = bf16_ref.bf16_fwd(q, k, v, False, tile_k=ops.bf16_fwd_key_step(384), mode="contract", lazy_tau=ops.BF16_RESCALE_TAU)
    assert (O.cpu() - Or).abs().max() < 2.5e-2 and (lse.cpu() - lser).abs().max() < 2e-2


@pytest.mark.parametrize("nsplit", [1, 2])
@pytest.mark.parametrize("S,D", [(200, 128), (333, 64), (65, 128), (130, 64)])
def test_jvp_ragged_matches_oracle(S, D, nsplit):
    from oracle import jvp_ref
    from quantizedattention_b200 import ops
    shape = (1, 2, S, D)
    g = torch.Generator().manual_seed(6100 + S + D)
    q, k, v, tq, tk, tv = [torch.randn(shape, generator=g) for _ in range(6)]
    O, tO, lse = ops.jvp_fwd(*[t.cuda() for t in (q, k, v, tq, tk, tv)], nsplit=nsplit)
    torch.cuda.synchronize()
    assert O.shape == shape and tO.shape == shape and lse.shape == (2, S)
    Oe, tOe, lsee = jvp_ref.jvp_fwd(q, k, v, tq, tk, tv, tile_k=128 if D == 64 else 64, operand_dtype=torch.bfloat16)
    assert (O.cpu() - Oe).abs().max() < 2e-3 and (tO.cpu() - tOe).abs().max() < 8e-3
    assert (lse.cpu() - lsee).abs().max() < 1e-3


def test_flash_atten_2_bf16_autograd_ragged():
    """The public entry point end to end on S = 500 against fp32 PyTorch attention + autograd."""
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_bf16 as A
    shape = (1, 2, 500, 128)
    g = torch.Generator().manual_seed(62)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    qh, kh, vh = q.half().cuda().requires_grad_(), k.half().cuda().requires_grad_(), v.bfloat16().cuda().requires_grad_()
    O = A.flash_atten_2_bf16(qh, kh, vh, True)
    O.backward(dO.cuda())
    qf, kf, vf = q.half().float().requires_grad_(), k.half().float().requires_grad_(), v.bfloat16().float().requires_grad_()
    Ob = baseline_pytorch_attention(qf, kf, vf, 128, True)
    Ob.backward(dO)
    assert (O.detach().cpu() - Ob.detach()).abs().max() < 3e-2
    for name, a, b in zip("qkv", (qh, kh, vh), (qf, kf, vf)):
        assert _rel(a.grad.cpu(), b.grad) < 2e-2, (name, _rel(a.grad.cpu(), b.grad))
