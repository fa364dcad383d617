"""Bit-exact parity of the int8 quantisation pre-pass against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + offset).to(torch.float16)


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("blk", [32, 64, 128, 256])
def test_quant_block_bit_exact(D, blk):
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    x = _mk((2, 3, 512, D), 100 + blk + D)
    x[0, 0, :blk] = 0                                   # all-zero block (LEDGER I-4)
    x[0, 1, :blk] *= 1e-3                               # subnormal fp16 scale
    x[1, 2, 5, 7] = 60000.0                             # near fp16 max
    x[1, 0, blk:2 * blk] *= 37.0
    qi, s = ops.quant_block(x.cuda(), blk)
    qo, so = int8_ref.quant_block(x.view(-1, D), blk)
    assert torch.equal(s.cpu(), so)
    assert torch.equal(qi.cpu(), qo)


@pytest.mark.parametrize("D", [64, 128])
def test_k_mean_and_smoothed_quant(D):
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    k = _mk((2, 4, 1024, D), 7 + D, offset=1.5)
    km = ops.k_mean(k.cuda())
    ref64 = k.double().mean(dim=2, keepdim=True)
    # fp32-accumulated mean rounded to fp16: allow one fp16 ulp against the fp64 mean
    ulp = torch.finfo(torch.float16).eps * ref64.abs().clamp_min(2.0 ** -14)
    assert ((km.cpu().double() - ref64).abs() <= ulp).all()
    # smoothing + quantisation: bit-exact given the SAME fp16 mean (contract stated at the kernel boundary)
    qi, s = ops.quant_block(k.cuda(), 128, mean=km, rows_per_head=1024)
    ks = int8_ref.smooth_k(k, km.cpu())
    qo, so = int8_ref.quant_block(ks.view(-1, D), 128)
    assert torch.equal(s.cpu(), so) and torch.equal(qi.cpu(), qo)


def test_quant_large_random_bit_exact():
    """Many values across magnitudes: exercises the Markstein-corrected division against IEEE fp16 divide."""
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(64 * 1024, 128, generator=g) * torch.exp(torch.randn(64 * 1024, 1, generator=g) * 2)).to(torch.float16)
    qi, s = ops.quant_block(x.cuda(), 32)
    qo, so = int8_ref.quant_block(x, 32)
    assert torch.equal(s.cpu(), so) and torch.equal(qi.cpu(), qo)
