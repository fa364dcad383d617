"""NVFP4 (microscaling) forward: bit-exact e2m1 codes, e4m3 block scales (in the tcgen05.cp atom layout) and
per-head scales against the eager definition (oracle/fp4_ref.py); O / lse of the kernel against the definition run on the
same quantised operands; quantisation-level agreement with fp32 attention math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a - b).abs().max().item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def _inputs(shape, seed, kind="randn"):
    g = torch.Generator().manual_seed(seed)
    q, k, v = [torch.randn(shape, generator=g) for _ in range(3)]
    if kind == "offset":                     # large per-channel offset on K (exercises the smoothing), logits scaled up
        k = k + 4.0 * torch.randn(1, shape[1], 1, shape[3], generator=g)
        q = q * 2.0
    if kind == "zeros":                      # all-zero 16-blocks, an all-zero V head and tiny magnitudes
        q[:, :, :, 16:32] = 0
        v[:, 0] = 0
        k = k * 1e-3
    if kind == "ties":                       # every quotient x / (sf * sg) exactly on an e2m1 rounding threshold, or one fp16 ulp beside it:
        t = torch.tensor([6, .25, .75, 1.25, 1.75, 2.5, 3.5, 5, -.25, -.75, -1.25, -1.75, -2.5, -3.5, -5, .5])
        q = (t / 64).repeat(shape[3] // 16).expand(shape).clone()          # head amax 42 -> sg = 2^-6; block amax 6 / 64 -> sf = 1
        q[:, :, 0, :16] = t * 7                                            # the block that carries the head amax (sf = 448)
        v = (t / 64).repeat(shape[2] // 16)[:, None].expand(shape).clone()  # V: blocks of 16 keys
        v[:, :, :16, 0] = t * 7
        q, v = q.to(torch.float16), v.to(torch.float16)
        q.view(torch.int16)[:, :, 1::3] += 1; q.view(torch.int16)[:, :, 2::3] -= 1          # magnitude one ulp up / down
        v.view(torch.int16)[..., 1::3] += 1; v.view(torch.int16)[..., 2::3] -= 1
        return [q, q.clone(), v]
    return [t.to(torch.float16) for t in (q, k, v)]


@pytest.mark.parametrize("shape,kind", [((1, 2, 256, 128), "randn"), ((2, 2, 512, 128), "offset"), ((1, 2, 384, 128), "zeros"),
                                        ((1, 2, 200, 128), "offset"), ((2, 1, 333, 128), "randn"),      # ragged: zero padding to 256 / 384
                                        ((1, 2, 256, 128), "ties"),           # the division fallback of the code conversion
                                        ((1, 3, 8192, 128), "offset"),        # V in one pass: the 64 CTAs of a head wait for each other
                                        ((1, 1, 20480, 128), "randn")])       # 160 tiles per head > 148 SMs: V falls back to two passes
def test_fp4_quantisation_is_bit_exact(shape, kind):
    from oracle import fp4_ref
    from quantizedattention_b200 import attention_fp4 as F
    q, k, v = _inputs(shape, 400 + shape[2], kind)
    o = F.quantise_fp4(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    ref = fp4_ref.quantise_inputs(q, k, v)
    assert torch.equal(o.k_mean.cpu().view(ref["k_mean"].shape), ref["k_mean"])
    for name in ("sgq", "sgk", "sgv"):
        assert torch.equal(getattr(o, name).cpu(), ref[name]), name
    for name in ("sfq", "sfk", "sfv"):
        assert torch.equal(getattr(o, name).cpu(), ref[name]), name
    for name in ("q4", "k4", "vt4"):
        assert torch.equal(getattr(o, name).cpu(), ref[name]), name


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("shape,kind", [((1, 2, 256, 128), "randn"), ((1, 4, 1024, 128), "randn"), ((2, 2, 512, 128), "offset"),
                                        ((1, 2, 384, 128), "zeros")])
def test_fp4_fwd_matches_definition_and_fp32_math(shape, kind, variant):
    from oracle import fp4_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp4 as F
    q, k, v = _inputs(shape, 500 + shape[2], kind)
    O, lse = F.fp4_fwd_prequant(F.quantise_fp4(q.cuda(), k.cuda(), v.cuda()), variant=variant)
    torch.cuda.synchronize()
    Oref, lse_ref, _ = fp4_ref.fp4_fwd(q, k, v, step=128 if variant == 0 else 64)
    fin = torch.isfinite(Oref.float())
    assert torch.isfinite(O.float().cpu()[fin]).all()
    mx, cos = _stats(O.cpu()[fin], Oref[fin])
    assert mx < 1e-2 and cos > 0.9995, (mx, cos)         # ex2.approx and reciprocal-multiply flip a few e2m1 roundings of P
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-3
    if kind != "zeros":
        base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], False)
        mx, cos = _stats(O.cpu(), base)
        assert cos > (0.95 if kind == "offset" else 0.975), (mx, cos)   # 4-bit operands: quantisation-level agreement only (sharper softmax: lower)


def test_sage_attention_3_fp4_validates():
    from quantizedattention_b200 import attention_fp4 as F
    q = torch.randn(1, 2, 256, 128, device="cuda", dtype=torch.float16)
    O = F.sage_attention_3_fp4(q, q, q)
    assert O.shape == q.shape and O.dtype == torch.float16 and not O.requires_grad
    with pytest.raises(TypeError):
        F.sage_attention_3_fp4(q.float(), q.float(), q.float())
    with pytest.raises(ValueError):
        F.sage_attention_3_fp4(q[..., :32].contiguous(), q[..., :32].contiguous(), q[..., :32].contiguous())


@pytest.mark.parametrize("variant", [0, 1])
def test_fp4_fwd_at_baseline_sequence_length(variant):
    """S = 8192, D = 128 (the metric's shape): every barrier phase, stage wrap and scale-factor tile of the benchmark run is
    exercised; quantised operands bit-exact, O / lse against the definition with the bars of the small shapes."""
    import os
    from oracle import fp4_ref
    from quantizedattention_b200 import attention_fp4 as F
    torch.set_num_threads(os.cpu_count() or 1)
    shape = (1, 2, 8192, 128)
    g = torch.Generator().manual_seed(8192 + variant)
    q, k, v = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(3)]
    k = (k.float() + 0.5).to(torch.float16)
    o = F.quantise_fp4(q.cuda(), k.cuda(), v.cuda())
    O, lse = F.fp4_fwd_prequant(o, variant=variant)
    O2, _ = F.fp4_fwd_prequant(o, variant=variant)
    torch.cuda.synchronize()
    assert torch.equal(O, O2)                                         # run-to-run bit equality
    Oref, lse_ref, ref = fp4_ref.fp4_fwd(q, k, v, step=128 if variant == 0 else 64)
    for name in ("q4", "k4", "vt4", "sfq", "sfk", "sfv", "sgq", "sgk", "sgv"):
        assert torch.equal(getattr(o, name).cpu(), ref[name]), name
    mx, cos = _stats(O.cpu(), Oref)
    assert mx < 1e-2 and cos > 0.9995, (mx, cos)
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("shape", [(1, 2, 256, 128), (2, 2, 1024, 128)])
def test_fp4_fwd_causal_matches_definition_and_baseline(shape):
    """Strict causal mask (key < query; row 0 = average over all keys), as on the int8 path."""
    from oracle import fp4_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp4 as F
    q, k, v = _inputs(shape, 700 + shape[2], "randn")
    O, lse = F.fp4_fwd_prequant(F.quantise_fp4(q.cuda(), k.cuda(), v.cuda()), causal=True)
    torch.cuda.synchronize()
    Oref, lse_ref, _ = fp4_ref.fp4_fwd(q, k, v, step=128, causal=True)
    assert torch.isfinite(O.float()).all()
    mx, cos = _stats(O.cpu(), Oref)
    assert mx < 2e-2 and cos > 0.9995, (mx, cos)             # early rows have few keys: single e2m1 flips weigh more
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-3
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], True)
    assert (O[:, :, 0].cpu().float() - base[:, :, 0]).abs().max().item() < 0.1      # row 0: mean of V, quantised
    mx, cos = _stats(O.cpu(), base)
    assert cos > 0.97, (mx, cos)
    with pytest.raises(RuntimeError):
        F.fp4_fwd_prequant(F.quantise_fp4(q.cuda(), k.cuda(), v.cuda()), variant=1, causal=True)


@pytest.mark.parametrize("S,causal", [(200, False), (333, True), (129, False), (1000, True)])
def test_fp4_fwd_ragged_sequence_lengths(S, causal):
    """Sequence lengths that are not multiples of 128 (the reference's hl.tile clamps its last tile): zero padding per head,
    padded keys masked in the kernel, padded K rows kept at zero by the quantiser."""
    from oracle import fp4_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp4 as F
    q, k, v = _inputs((1, 2, S, 128), 900 + S, "offset")
    O = F.sage_attention_3_fp4(q.cuda(), k.cuda(), v.cuda(), causal=causal)
    torch.cuda.synchronize()
    assert O.shape == (1, 2, S, 128)
    Oref, _, _ = fp4_ref.fp4_fwd(q, k, v, step=128, causal=causal)
    mx, cos = _stats(O.cpu(), Oref)
    assert torch.isfinite(O.float()).all() and mx < 2e-2 and cos > 0.9995, (mx, cos)
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), 128, causal)
    assert _stats(O.cpu(), base)[1] > 0.95


@pytest.mark.parametrize("shape,causal", [((1, 8, 1024, 64), False), ((2, 3, 300, 64), True)])
def test_fp4_fwd_head_dim_64(shape, causal):
    """The reference's D = 64 shapes (configs[0]: B=1 H=8 S=1024 D=64) run with zero-padded head columns and sm_scale = 1/8."""
    from oracle import fp4_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_fp4 as F
    q, k, v = _inputs(shape, 64 + shape[2], "randn")
    O = F.sage_attention_3_fp4(q.cuda(), k.cuda(), v.cuda(), causal=causal)
    torch.cuda.synchronize()
    assert O.shape == shape
    Oref, _, _ = fp4_ref.fp4_fwd(q, k, v, step=128, causal=causal)
    mx, cos = _stats(O.cpu(), Oref)
    assert mx < 2e-2 and cos > 0.9995, (mx, cos)
    assert _stats(O.cpu(), baseline_pytorch_attention(q.float(), k.float(), v.float(), 64, causal))[1] > 0.97
