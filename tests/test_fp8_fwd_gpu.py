"""fp8 (e4m3) forward: bit-exact e4m3 codes and scales against the eager definition (oracle/fp8_ref.py),
O / lse within the int8 path's bars of it, and quantisation-level agreement with fp32 attention math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a - b).abs().max().item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


@pytest.mark.parametrize("shape", [(1, 2, 256, 128), (1, 8, 1024, 64), (2, 2, 512, 128)])
def test_fp8_fwd_matches_definition_and_fp32_math(shape):
    from oracle import fp8_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp8 as F
    g = torch.Generator().manual_seed(800 + shape[2] + shape[3])
    q, k, v = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(3)]
    out = F.helion_atten_fp8_fwd(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    Oref, lse_ref, (qb, kb, vb, sq, sk, sv) = fp8_ref.fp8_fwd(q, k, v)
    assert out[2].dtype == torch.float8_e4m3fn
    assert torch.equal(out[2].view(torch.uint8).cpu(), qb) and torch.equal(out[4].view(torch.uint8).cpu(), vb)
    assert torch.equal(out[3].t().contiguous().view(torch.uint8).cpu(), kb)
    assert torch.equal(out[5].cpu(), sq) and torch.equal(out[6].cpu(), sk) and torch.equal(out[7].cpu(), sv)
    mx, cos = _stats(out[0].cpu(), Oref)
    assert mx < 8e-3 and cos > 0.9999, (mx, cos)             # e4m3 P has a 3-bit mantissa: RN ties differ more often than int8
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], False)
    mx, cos = _stats(out[0].cpu(), base)
    assert mx < 0.15 and cos > 0.995, (mx, cos)              # e4m3 (3-bit mantissa) quantisation noise


def test_sage_attention_3_fp8_smooths_k_and_validates():
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp8 as F
    shape = (1, 4, 512, 128)
    g = torch.Generator().manual_seed(9)
    q, k, v = [torch.randn(shape, generator=g) for _ in range(3)]
    k = k + 2.0                                              # large channel offset: smoothing keeps the e4m3 range for the signal
    O = F.sage_attention_3_fp8(q.half().cuda(), k.half().cuda(), v.half().cuda())
    assert O.dtype == torch.float16 and not O.requires_grad
    base = baseline_pytorch_attention(q.half().float(), k.half().float(), v.half().float(), 128, False)
    mx, cos = _stats(O.cpu(), base)
    assert cos > 0.995, (mx, cos)
    with pytest.raises(ValueError):
        F.sage_attention_3_fp8(q[:, :, :100].half().cuda(), k[:, :, :100].half().cuda(), v[:, :, :100].half().cuda())
    with pytest.raises(TypeError):
        F.sage_attention_3_fp8(q.cuda(), k.cuda(), v.cuda())
