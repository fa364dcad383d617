"""Parity of the bf16 bias-corrected forward (row a5) and the JVP kernel (row a8) against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a - b).abs().max().item(), ((a - b) ** 2).mean().item()


@pytest.mark.parametrize("tau", [0.0, 8.0])
@pytest.mark.parametrize("nsplit", [1, 2, 3])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 2, 256, 128), (2, 2, 512, 64), (1, 1, 128, 128), (1, 2, 1024, 128), (1, 1, 2048, 64)])
def test_bf16_fwd_matches_oracle(shape, causal, nsplit, tau):
    from oracle import bf16_ref
    from oracle.baseline import baseline_lse_log2, baseline_pytorch_attention
    from quantizedattention_b200 import ops
    g = torch.Generator().manual_seed(3000 + shape[2] + shape[3] + causal)
    q, k, v = [torch.randn(shape, generator=g) for _ in range(3)]
    q, k, v = q.half(), k.half(), v.bfloat16()
    O, lse = ops.bf16_fwd(q.cuda(), k.cuda(), v.cuda(), causal, nsplit=nsplit, rescale_tau=tau)
    torch.cuda.synchronize()
    assert O.dtype == torch.float32 and lse.shape == (shape[0] * shape[1], shape[2])
    Or, lser = bf16_ref.bf16_fwd(q, k, v, causal, tile_k=ops.bf16_fwd_key_step(shape[2], nsplit), mode="contract", lazy_tau=tau)
    mx, mse = _stats(O.cpu(), Or)
    assert mx < 2.5e-2 and mse < 5e-6, (mx, mse)             # bf16 P rounding class: reference yardstick max-abs 2e-2, MSE 3e-6
    assert (lse.cpu() - lser).abs().max() < 2e-2
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], causal)
    mx, mse = _stats(O.cpu(), base)
    assert mx < 3e-2 and mse < 5e-6, (mx, mse)
    # lse against fp32 math (row 0 of a causal head is the documented special case)
    lb = baseline_lse_log2(q, k, causal).reshape(lse.shape)
    sl = slice(1, None) if causal else slice(None)
    assert (lse.cpu()[:, sl] - lb[:, sl]).abs().max() < 2e-2
    if causal:      # LEDGER B-1: row 0 = mean over ALL keys
        assert (O.cpu()[:, :, 0] - v.float().mean(dim=2)).abs().max() < 1e-5


def test_bf16_fwd_stress_duplicate_keys():
    """Rows with duplicated keys -> repeated row max -> exercises the bias-correction branch (m' doubled)."""
    from oracle import bf16_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import ops
    g = torch.Generator().manual_seed(9)
    q, k, v = [torch.randn(1, 2, 256, 64, generator=g) for _ in range(3)]
    k[:, :, 1::2] = k[:, :, 0::2]                               # every key appears twice
    q = q * 3.0
    q, k, v = q.half(), k.half(), v.bfloat16()
    O, _ = ops.bf16_fwd(q.cuda(), k.cuda(), v.cuda(), False)
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), 64, False)
    assert torch.isfinite(O).all()
    Or, _ = bf16_ref.bf16_fwd(q, k, v, False, tile_k=ops.bf16_fwd_key_step(256), mode="contract", lazy_tau=ops.BF16_RESCALE_TAU)
    assert (O.cpu() - Or).abs().max() < 2.5e-2               # same algorithm: only MMA summation order differs
    assert (O.cpu() - base).abs().max() < 1e-1               # bf16 logits at |s| ~ 4..8: the oracle itself is 6e-2 off


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("nsplit", [1, 2])
@pytest.mark.parametrize("ones", [True, False])
def test_jvp_matches_oracle_and_torch_func(ones, nsplit, D):
    from oracle import jvp_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import ops
    shape = (2, 2, 256, D)
    g = torch.Generator().manual_seed(41 + ones)
    q, k, v = [torch.randn(shape, generator=g) for _ in range(3)]
    tq, tk, tv = [torch.ones(shape) if ones else torch.randn(shape, generator=g) for _ in range(3)]
    O, tO, lse = ops.jvp_fwd(*[t.cuda() for t in (q, k, v, tq, tk, tv)], nsplit=nsplit)
    torch.cuda.synchronize()
    # same-operand-rounding oracle (bf16 MMA operands, fp32 accumulate): tight
    Oe, tOe, lsee = jvp_ref.jvp_fwd(q, k, v, tq, tk, tv, tile_k=128 if D == 64 else 64, operand_dtype=torch.bfloat16)
    assert (O.cpu() - Oe).abs().max() < 2e-3 and (tO.cpu() - tOe).abs().max() < 8e-3
    assert (lse.cpu() - lsee).abs().max() < 1e-3
    # fp32 truth: torch.func.jvp of the reference baseline (attention_jvp.py:254-258); yardstick atol 1e-2
    Ob, tOb = torch.func.jvp(baseline_pytorch_attention, (q, k, v), (tq, tk, tv))
    mxO, mseO = _stats(O.cpu(), Ob)
    mxT, mseT = _stats(tO.cpu(), tOb)
    assert mxO < 1e-2 and mseO < 5e-7, (mxO, mseO)
    assert mxT < (1e-2 if ones else 4e-2) and mseT < (1e-6 if ones else 2e-5), (mxT, mseT)


def test_jvp_forward_mode_ad_integration():
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_jvp as J
    shape = (1, 2, 128, 64)
    g = torch.Generator().manual_seed(5)
    q, k, v, tq, tk, tv = [torch.randn(shape, generator=g) for _ in range(6)]
    O, tO = torch.func.jvp(J.jvp_attention, tuple(t.cuda() for t in (q, k, v)), tuple(t.cuda() for t in (tq, tk, tv)))
    Ob, tOb = torch.func.jvp(baseline_pytorch_attention, (q, k, v), (tq, tk, tv))
    assert (O.cpu() - Ob).abs().max() < 1e-2 and (tO.cpu() - tOb).abs().max() < 4e-2
    # raw reference-named callable
    O2, tO2, lse = J.helion_attention_jvp_forward_fp32(*[t.cuda() for t in (q, k, v, tq, tk, tv)])
    assert torch.equal(O2, O) and torch.equal(tO2, tO) and lse.shape == (2, 128)


def test_jvp_dual_level_forward_ad():
    """torch.autograd.forward_ad (the README's 'torch forward mode auto-differentiation', README.md:19-22)."""
    import torch.autograd.forward_ad as fwAD
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_jvp as J
    shape = (1, 1, 128, 64)
    g = torch.Generator().manual_seed(6)
    q, k, v, tq, tk, tv = [torch.randn(shape, generator=g) for _ in range(6)]
    with fwAD.dual_level():
        dq, dk, dv = [fwAD.make_dual(a.cuda(), b.cuda()) for a, b in ((q, tq), (k, tk), (v, tv))]
        out = J.jvp_attention(dq, dk, dv)
        O, tO = fwAD.unpack_dual(out)
    Ob, tOb = torch.func.jvp(baseline_pytorch_attention, (q, k, v), (tq, tk, tv))
    assert (O.cpu() - Ob).abs().max() < 1e-2 and (tO.cpu() - tOb).abs().max() < 4e-2
