"""Host-staged pipeline (quantizedattention_b200/host_pipeline.py): chunking over batch x head and overlapping the
copies must not change results - every (b, h) is an independent problem."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,hc", [((2, 4, 256, 64), 2), ((1, 6, 384, 128), 4), ((1, 2, 128, 128), 32), ((2, 8, 256, 64), 4)])
def test_host_pipeline_matches_device_api(shape, hc):
    from quantizedattention_b200 import attention_int8 as A
    from quantizedattention_b200.host_pipeline import HostStagedSageAttention
    g = torch.Generator().manual_seed(4242 + shape[2])
    q, k, v, dO = [torch.randn(shape, generator=g).half().pin_memory() for _ in range(4)]
    pipe = HostStagedSageAttention(heads_per_chunk=hc)
    for _ in range(2):                                              # second call reuses the staging buffers
        O, dq, dk, dv = pipe(q, k, v, dO)
        torch.cuda.synchronize()
    qr, kr, vr = [t.cuda().requires_grad_() for t in (q, k, v)]
    Or = A.sage_attention_3_int8(qr, kr, vr)
    Or.backward(dO.cuda())
    torch.cuda.synchronize()
    assert O.is_pinned() and not O.is_cuda
    assert torch.equal(O, Or.detach().cpu())                        # forward is deterministic per head
    assert torch.equal(dk, kr.grad.cpu()) and torch.equal(dv, vr.grad.cpu())
    # dQ is reduced over k-tiles with fp32 atomics (order not fixed): equal up to fp32 summation order, then fp16
    assert (dq.float() - qr.grad.cpu().float()).abs().max() <= 2e-3 * qr.grad.abs().max().item() + 1e-6


def test_host_pipeline_forward_only_and_errors():
    from quantizedattention_b200 import attention_int8 as A
    from quantizedattention_b200.host_pipeline import sage_attention_3_int8_host
    g = torch.Generator().manual_seed(7)
    q, k, v = [torch.randn(1, 4, 256, 64, generator=g).half().pin_memory() for _ in range(3)]
    O = sage_attention_3_int8_host(q, k, v, heads_per_chunk=3)      # 3 does not divide 4: falls back to 2 per chunk
    torch.cuda.synchronize()
    with torch.no_grad():
        Or = A.sage_attention_3_int8(q.cuda(), k.cuda(), v.cuda())
    assert torch.equal(O, Or.cpu())
    with pytest.raises(RuntimeError):
        sage_attention_3_int8_host(q.clone(), k, v)                  # pageable host memory
    with pytest.raises(TypeError):
        sage_attention_3_int8_host(q.cuda(), k, v)
