"""Ring-KV state continuation on the GPU: the K/V shards of a 2-rank ring are visited sequentially on one device
(no kernels wait on each other) and compared with the single-shot kernel and with the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,Bkv", [(64, 128), (128, 128), (128, 64), (64, 32)])
def test_ring_state_continuation_matches_single_shot(D, Bkv):
    from oracle import int8_ref
    from quantizedattention_b200 import ops
    from quantizedattention_b200.parallel import CudaRingKernels
    g = torch.Generator().manual_seed(17 + D)
    B, H, S = 1, 2, 512
    q, k, v = [torch.randn(B, H, S, D, generator=g).to(torch.float16) for _ in range(3)]
    kern = CudaRingKernels()
    world, Sl, BH = 2, S // 2, B * H
    km_sum = sum(kern.token_sum(k[:, :, r * Sl:(r + 1) * Sl].contiguous().cuda()) for r in range(world))
    km = (km_sum / S).to(torch.float16)
    shards = []
    for r in range(world):
        sl = slice(r * Sl, (r + 1) * Sl)
        q_i8, sq = kern.quant(q[:, :, sl].contiguous().cuda(), 128)
        k_i8, sk = kern.quant(k[:, :, sl].contiguous().cuda(), Bkv, mean=km, rows_per_head=Sl)
        v_i8, sv = kern.quant(v[:, :, sl].contiguous().cuda(), Bkv)
        shards.append((q_i8, sq, (k_i8, v_i8, sk, sv)))
    outs = []
    for r in range(world):                      # rank r visits its own shard first, then the one received from r-1
        q_i8, sq, _ = shards[r]
        st = kern.attend(q_i8, sq, shards[r][2], None, BH, Sl, Sl, D, 128, Bkv, last=False)
        O, lse16, lse32 = kern.attend(q_i8, sq, shards[(r - 1) % world][2], st, BH, Sl, Sl, D, 128, Bkv, last=True)
        outs.append(O.view(B, H, Sl, D).cpu())
    O_ring = torch.cat(outs, dim=2)
    full = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, km.cpu()), v, 128, Bkv, per_head=True)
    assert (O_ring.float() - full[0].float()).abs().max() < 6e-3       # different k-tile order: tolerance, not bitwise
    # oracle ring step (same order) is tighter
    c = lambda t: t.cpu()
    r = 1
    q_i8, sq, _ = shards[r]
    st = int8_ref.int8_attend_state(c(q_i8), c(sq), *[c(t) for t in shards[1][2]], None, BH, Sl, Sl, D, 128, Bkv, False)
    Oo, _, _ = int8_ref.int8_attend_state(c(q_i8), c(sq), *[c(t) for t in shards[0][2]], st, BH, Sl, Sl, D, 128, Bkv, True)
    assert (outs[1].reshape(-1, D).float() - Oo.float()).abs().max() < 5e-3


@pytest.mark.parametrize("D", [64, 128])
def test_causal_chunk_continuation_matches_single_shot(D):
    """The building blocks of the zig-zag causal ring on one device (world 1: the rank owns chunks 0 and 1 = the whole
    sequence): diagonal chunk with state out, then the late query chunk continues over the early K/V chunk; against the
    single-shot causal kernel's oracle."""
    from oracle import int8_ref
    from quantizedattention_b200.parallel import ring_int8_attention_fwd_causal
    g = torch.Generator().manual_seed(23 + D)
    B, H, S = 1, 2, 512
    q, k, v = [torch.randn(B, H, S, D, generator=g).to(torch.float16) for _ in range(3)]
    k = (k.float() + 0.5).to(torch.float16)
    O, lse, km = ring_int8_attention_fwd_causal(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    full = int8_ref.sage_forward(q, k, v, 128, 128, causal=True)
    assert torch.equal(km.cpu(), full[2])
    assert (O.cpu().float() - full[0].float()).abs().max() < 6e-3
    ref_lse = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, full[2]), v, 128, 128, return_lse32=True, causal=True)[10].view(B * H, S)
    assert (lse.cpu() - ref_lse).abs().max() < 3e-2
