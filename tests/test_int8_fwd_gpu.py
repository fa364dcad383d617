"""Parity of the fused int8 forward against the oracle and fp32 attention math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g).to(torch.float16) for _ in range(3)]


def _stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a - b).abs().max().item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


@pytest.mark.parametrize("nsplit", [0, 1, 2])
@pytest.mark.parametrize("shape,Bq", [((1, 2, 256, 128), 128), ((1, 8, 1024, 64), 128), ((2, 2, 512, 128), 32),
                                      ((1, 1, 128, 64), 64)])
def test_int8_fwd_matches_oracle(shape, Bq, nsplit):
    from oracle import int8_ref
    from oracle.baseline import baseline_pytorch_attention
    from quantizedattention_b200 import attention_int8 as A
    q, k, v = _mk(shape, 1000 + shape[2] + shape[3])
    A.set_block_sizes(Bq, 128)
    A._CFG["nsplit"] = nsplit
    try:
        out = A.helion_atten_int8_hl_dot_fwd(q.cuda(), k.cuda(), v.cuda(), _want_lse32=True)
    finally:
        A.set_block_sizes(128, 128)
        A._CFG["nsplit"] = 0
    torch.cuda.synchronize()
    ref = int8_ref.int8_fwd(q, k, v, Bq, 128, per_head=True, return_lse32=True)
    # quantised tensors and scales: bit-exact
    for i in (2, 3, 4, 5, 6, 7):
        assert torch.equal(out[i].cpu(), ref[i]), f"slot {i}"
    assert out[8] == Bq and out[9] == 128
    # O / lse: tolerance vs the eager oracle (same algorithm) and vs fp32 math
    mx, cos = _stats(out[0].cpu(), ref[0])
    assert mx < 5e-3 and cos > 0.99999, (mx, cos)
    assert (out[1].cpu().float() - ref[1].float()).abs().max() < 4e-2          # fp16 lse: 1-2 ulp at |lse| ~ 8..16
    assert (out[10].cpu() - ref[10]).abs().max() < 2e-3
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], False)
    mx, cos = _stats(out[0].cpu(), base)
    assert mx < 8e-2 and cos > 0.999, (mx, cos)                               # reference yardstick: atol 1e-2 class noise


@pytest.mark.parametrize("shape,Bq,Bkv", [((1, 2, 256, 128), 32, 32), ((2, 2, 256, 64), 32, 32), ((1, 2, 512, 128), 64, 64),
                                          ((1, 1, 256, 64), 128, 64), ((1, 2, 256, 128), 256, 32),
                                          ((1, 2, 512, 128), 128, 256), ((2, 1, 512, 64), 64, 256), ((1, 2, 768, 128), 256, 256)])
def test_int8_fwd_reference_default_block_sizes(shape, Bq, Bkv):
    """Bkv = 32 / 64: the reference's untuned default tunables (PowerOfTwoFragment(32, 256, 32), attention_int8.py:155-158)."""
    from oracle import int8_ref
    from quantizedattention_b200 import attention_int8 as A
    q, k, v = _mk(shape, 77 + Bq + Bkv + shape[3])
    A.set_block_sizes(Bq, Bkv)
    try:
        out = A.helion_atten_int8_hl_dot_fwd(q.cuda(), k.cuda(), v.cuda(), _want_lse32=True)
    finally:
        A.set_block_sizes(128, 128)
    ref = int8_ref.int8_fwd(q, k, v, Bq, Bkv, per_head=True, return_lse32=True)
    for i in (2, 3, 4, 5, 6, 7):
        assert torch.equal(out[i].cpu(), ref[i]), f"slot {i}"
    assert out[8:10] == (Bq, Bkv)
    mx, cos = _stats(out[0].cpu(), ref[0])
    assert mx < 5e-3 and cos > 0.99999, (mx, cos)
    assert (out[10].cpu() - ref[10]).abs().max() < 2e-3


def test_int8_fwd_against_real_reference_fixture(golden_dir):
    """The fixture holds the outputs of the UNMODIFIED reference (Bq = Bkv = 32, attention over the flattened B*H*S
    axis, LEDGER I-2).  The literal result is the per-head kernel on x.view(1, 1, B*H*S, D)."""
    import os
    from quantizedattention_b200 import attention_int8 as A
    fx = torch.load(os.path.join(golden_dir, "int8_B1H2S128D64_bq32_bkv32.pt"), weights_only=True)
    q, k, v = [fx[n].reshape(1, 1, -1, 64).cuda() for n in ("q", "k", "v")]
    A.set_block_sizes(32, 32)
    try:
        out = A.helion_atten_int8_hl_dot_fwd(q, k, v)
    finally:
        A.set_block_sizes(128, 128)
    O_ref, lse_ref, q_i8, k_i8_T, v_i8, sq, sk, sv = fx["fwd"]
    assert torch.equal(out[2].cpu(), q_i8) and torch.equal(out[3].cpu(), k_i8_T) and torch.equal(out[4].cpu(), v_i8)
    assert torch.equal(out[5].cpu(), sq) and torch.equal(out[6].cpu(), sk) and torch.equal(out[7].cpu(), sv)
    assert (out[0].cpu().float().reshape(-1) - O_ref.float().reshape(-1)).abs().max() < 5e-3
    assert (out[1].cpu().float() - lse_ref.float()).abs().max() < 4e-2


@pytest.mark.parametrize("ring", [False, True])
def test_int8_fwd_running_max_jump_beyond_fp32_range(ring):
    """A late key whose logit exceeds every earlier one by more than 126 log2 units: the rescale factor exp2(m - m') is
    exactly 0 in the kernel (ex2.approx.ftz), so the O accumulator must drop everything accumulated so far, consistently
    with l and lse (the lazy rescale used to skip the update).  Plain path and ring-state continuation."""
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    B, H, S, D = 1, 2, 256, 128
    g = torch.Generator().manual_seed(5)
    q = torch.full((B, H, S, D), 3.0).to(torch.float16)
    k = (0.01 * torch.randn(B, H, S, D, generator=g)).to(torch.float16)
    k[:, :, 128 + 17] = 3.0                                      # one key in the LAST tile: logit ~ 147 in the log2 domain
    v = torch.randn(B, H, S, D, generator=g).to(torch.float16)
    ref = int8_ref.int8_fwd(q, k, v, 128, 128, per_head=True, return_lse32=True)
    qi, sq = ops.quant_block(q.cuda(), 128); ki, sk = ops.quant_block(k.cuda(), 128); vi, sv = ops.quant_block(v.cuda(), 128)
    BH = B * H
    if not ring:
        O, _, lse32 = ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, 128, 128)
    else:                                                        # two K/V shards of one tile each, state carried between them
        kv = lambda t, r: t.view(BH, S, D)[:, r * 128:(r + 1) * 128].contiguous().view(-1, D)
        sc = lambda t, r: t.view(BH, 2)[:, r].contiguous()
        st = ops.int8_fwd_prequant(qi, kv(ki, 0), kv(vi, 0), sq, sc(sk, 0), sc(sv, 0), BH, S, 128, D, 128, 128, ring_state=True)
        O, _, lse32 = ops.int8_fwd_prequant(qi, kv(ki, 1), kv(vi, 1), sq, sc(sk, 1), sc(sv, 1), BH, S, 128, D, 128, 128, state_in=st)
    torch.cuda.synchronize()
    # every row attends (numerically only) to the one large key: O = its de-quantised V row
    assert torch.isfinite(O.float()).all()
    assert (O.cpu().float() - ref[0].reshape(-1, D).float()).abs().max() < 5e-3
    assert (lse32.cpu() - ref[10]).abs().max() < 2e-3
    expect = v[:, :, 128 + 17].float().reshape(BH, 1, D).expand(BH, S, D).reshape(-1, D)
    assert (O.cpu().float() - expect).abs().max() < 5e-2
