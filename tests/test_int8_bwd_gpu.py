"""Parity of the fused int8 backward against the contract oracle and fp32 autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("kernel", ["warp_specialised", "eight_warp"])
@pytest.mark.parametrize("shape", [(1, 2, 256, 128), (2, 2, 512, 64), (1, 1, 128, 128)])
def test_int8_bwd_matches_contract_oracle(shape, kernel):
    from oracle import int8_ref
    # the default kernel is the warp-specialised one; the 8-warp kernel stays selectable per call (QA_FLAG_BWD_8WARP)
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = shape
    g = torch.Generator().manual_seed(2000 + S + D)
    q, k, v, dO = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(4)]
    k = (k.float() + 1.0).to(torch.float16)                 # non-zero token mean: exercises the k_mean term
    out = A.SageAttention3_Int8_autograd_function.forward(q.cuda(), k.cuda(), v.cuda())
    O, lse16, kmean, q_i8, k_i8_T, v_i8, sq, sk, sv, Bq, Bkv = out
    # oracle on the SAME saved tensors (so only the backward is compared)
    c = lambda t: t.cpu()
    ref = int8_ref.int8_bwd_contract(dO, c(q_i8), c(sq), c(k_i8_T), c(kmean), c(sk), c(v_i8), c(sv), c(O), c(lse16), Bq, Bkv)
    got = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv,
                                         kernel="ws" if kernel == "warp_specialised" else "8warp")
    torch.cuda.synchronize()
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert _cos(a.cpu(), b) > 0.9995 and _rel(a.cpu(), b) < 3e-2, (name, _cos(a.cpu(), b), _rel(a.cpu(), b))


def test_sage_attention_autograd_end_to_end():
    """sage_attention_3_int8(...).backward() against fp32 PyTorch attention + autograd (the reference's own
    comparison, attention_int8.py:517-528, with dO ~ N(0,1) instead of the vacuous mse_loss gradients)."""
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    shape = (2, 4, 512, 128)
    g = torch.Generator().manual_seed(77)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    qh, kh, vh = [t.to(torch.float16).cuda().requires_grad_() for t in (q, k, v)]
    O = A.sage_attention_3_int8(qh, kh, vh)
    assert O.dtype == torch.float16 and O.shape == shape
    saved = O.grad_fn.saved_tensors
    assert saved[-1] is not None and saved[-1].dtype == torch.float32      # fp32 lse reached the backward (LEDGER I-15)
    O.backward(dO.to(torch.float16).cuda())
    qf, kf, vf = [t.to(torch.float16).float().requires_grad_() for t in (q, k, v)]
    Ob = baseline_pytorch_attention(qf, kf, vf, shape[3], False)
    Ob.backward(dO.to(torch.float16).float())
    assert (O.detach().cpu().float() - Ob.detach()).abs().max() < 8e-2
    for name, a, b in zip("qkv", (qh, kh, vh), (qf, kf, vf)):
        assert a.grad.dtype == torch.float16
        cs, rl = _cos(a.grad.cpu(), b.grad), _rel(a.grad.cpu(), b.grad)
        assert cs > 0.99 and rl < 0.2, (name, cs, rl)      # survey probe (128/128 blocks): cos 0.994-0.997, rel 0.12-0.15


@pytest.mark.parametrize("Bq,Bkv", [(32, 32), (64, 64), (32, 64), (64, 32), (128, 32), (32, 128), (64, 128)])
@pytest.mark.parametrize("shape", [(1, 2, 256, 128), (2, 1, 384, 64)])
def test_int8_bwd_reference_tunables(shape, Bq, Bkv):
    """Bq / Bkv below the MMA tile (the reference's PowerOfTwoFragment(32, 256, 32), attention_int8.py:155-158, forwarded
    to the backward :65,81,92; 32/32 is its untuned default): per-[Bq,Bkv]-block P / dS scales, per-Bq dO and Q scales."""
    from oracle import int8_ref
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = shape
    g = torch.Generator().manual_seed(3000 + S + D + Bq + 7 * Bkv)
    q, k, v, dO = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(4)]
    k = (k.float() + 1.0).to(torch.float16)
    out = A.SageAttention3_Int8_autograd_function.apply(q.cuda(), k.cuda(), v.cuda(), Bq=Bq, Bkv=Bkv)
    O, lse16, kmean, q_i8, k_i8_T, v_i8, sq, sk, sv, bq, bkv = out
    assert (bq, bkv) == (Bq, Bkv)
    c = lambda t: t.cpu()
    ref = int8_ref.int8_bwd_contract(dO, c(q_i8), c(sq), c(k_i8_T), c(kmean), c(sk), c(v_i8), c(sv), c(O), c(lse16), Bq, Bkv)
    got = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv)
    torch.cuda.synchronize()
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert _cos(a.cpu(), b) > 0.9995 and _rel(a.cpu(), b) < 3e-2, (name, _cos(a.cpu(), b), _rel(a.cpu(), b))


def test_int8_default_tunables_train_end_to_end():
    """`set_block_sizes(32, 32)` then `.backward()` (VERDICT r01 missing 6: used to raise)."""
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    shape = (1, 2, 512, 128)
    g = torch.Generator().manual_seed(78)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    qh, kh, vh = [t.to(torch.float16).cuda().requires_grad_() for t in (q, k, v)]
    A.sage_attention_3_int8(qh, kh, vh, Bq=32, Bkv=32).backward(dO.to(torch.float16).cuda())
    qf, kf, vf = [t.to(torch.float16).float().requires_grad_() for t in (q, k, v)]
    baseline_pytorch_attention(qf, kf, vf, shape[3], False).backward(dO.to(torch.float16).float())
    for name, a, b in zip("qkv", (qh, kh, vh), (qf, kf, vf)):
        cs, rl = _cos(a.grad.cpu(), b.grad), _rel(a.grad.cpu(), b.grad)
        assert cs > 0.995 and rl < 0.13, (name, cs, rl)     # survey probe at 32/32: cos 0.9985-0.9991, rel 0.07-0.09


@pytest.mark.parametrize("shape,Bq,Bkv", [((1, 2, 256, 128), 128, 128), ((2, 2, 512, 64), 128, 128), ((1, 2, 256, 128), 32, 32)])
def test_int8_bwd_sagebwd_option(shape, Bq, Bkv):
    """dP = dO V^T kept in fp16 (SageAttention3's SageBwd) - against the oracle with the same option, and
    closer to fp32 autograd than the fully quantised backward."""
    from oracle import int8_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    B, H, S, D = shape
    g = torch.Generator().manual_seed(3500 + S + D + Bq)
    q, k, v, dO = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(4)]
    out = A.SageAttention3_Int8_autograd_function.apply(q.cuda(), k.cuda(), v.cuda(), Bq=Bq, Bkv=Bkv)
    O, lse16, kmean, q_i8, k_i8_T, v_i8, sq, sk, sv, bq, bkv = out
    c = lambda t: t.cpu()
    ref = int8_ref.int8_bwd_contract(dO, c(q_i8), c(sq), c(k_i8_T), c(kmean), c(sk), c(v_i8), c(sv), c(O), c(lse16), Bq, Bkv, v_fp16=v)
    got = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv, v_fp16=v.cuda())
    torch.cuda.synchronize()
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert _cos(a.cpu(), b) > 0.9995 and _rel(a.cpu(), b) < 3e-2, (name, _cos(a.cpu(), b), _rel(a.cpu(), b))
    # accuracy: the option must not be worse than the fully quantised backward against fp32 autograd
    qf, kf, vf = [t.float().requires_grad_() for t in (q, k, v)]
    baseline_pytorch_attention(qf, kf, vf, D, False).backward(dO.float())
    plain = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv)
    for name, a, b, r in zip(("dq", "dk"), got[:2], plain[:2], (qf.grad, kf.grad)):
        assert _rel(a.cpu(), r) <= _rel(b.cpu(), r) * 1.02, (name, _rel(a.cpu(), r), _rel(b.cpu(), r))
    # through autograd
    qh, kh, vh = [t.cuda().requires_grad_() for t in (q, k, v)]
    A.sage_attention_3_int8(qh, kh, vh, Bq=Bq, Bkv=Bkv, sage_bwd=True).backward(dO.cuda())
    # (the autograd path recomputes P from the fp32 lse, the direct call above from the fp16 one the reference returns)
    assert _rel(kh.grad, got[1]) < 2e-2 and _rel(vh.grad, got[2]) < 2e-2 and _rel(qh.grad, got[0]) < 2e-2
