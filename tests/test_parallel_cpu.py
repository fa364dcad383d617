"""Host logic of the multi-GPU partitioning, exercised with world_size = 2 on CPU (gloo).  The device operations are
bound to the oracle (tests may use it); the GPU box runs the same ring with the CUDA kernels (test_parallel_gpu.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleRingKernels:
    def token_sum(self, k):
        return k.float().sum(dim=2, keepdim=True)

    def quant(self, x, blk, mean=None, rows_per_head=None):
        from oracle import int8_ref
        if mean is not None:
            x = int8_ref.smooth_k(x, mean)
        return int8_ref.quant_block(x.reshape(-1, x.shape[-1]), blk)

    def attend(self, q_i8, sq, kv, state, BH, Sq, Sk, D, Bq, Bkv, last):
        from oracle import int8_ref
        k_i8, v_i8, sk, sv = kv
        return int8_ref.int8_attend_state(q_i8, sq, k_i8, v_i8, sk, sv, state, BH, Sq, Sk, D, Bq, Bkv, last)


    def attend_causal_diag(self, q_i8, sq, kv, BH, S, D):
        from oracle import int8_ref
        k_i8, v_i8, sk, sv = kv
        return int8_ref.int8_attend_state(q_i8, sq, k_i8, v_i8, sk, sv, None, BH, S, S, D, 128, 128, False, causal_diag=True)

    def v_token_sum(self, v_i8, sv, BH, S, D, Bkv):
        return (v_i8.view(BH, S // Bkv, Bkv, D).float() * sv.view(BH, S // Bkv, 1, 1).float()).sum(dim=(1, 2))


def _full_inputs():
    g = torch.Generator().manual_seed(123)
    q, k, v = [torch.randn(1, 2, 512, 64, generator=g).to(torch.float16) for _ in range(3)]
    k = (k.float() + 0.5).to(torch.float16)
    return q, k, v


def _ring_worker(rank, world, port, ret, Bq=128, Bkv=128):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from quantizedattention_b200.parallel import ring_int8_attention_fwd
    q, k, v = _full_inputs()
    Sl = q.shape[2] // world
    sl = slice(rank * Sl, (rank + 1) * Sl)
    O, lse16, lse32, km = ring_int8_attention_fwd(q[:, :, sl].contiguous(), k[:, :, sl].contiguous(), v[:, :, sl].contiguous(),
                                                  Bq, Bkv, kernels=OracleRingKernels())
    ret[rank] = (O, lse32, km)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("Bq,Bkv", [(128, 128), (32, 64)])
def test_ring_kv_world2_matches_single_device(Bq, Bkv):
    from oracle import int8_ref
    world, port = 2, 29500 + (os.getpid() % 500) + Bkv
    ret = mp.Manager().dict()
    mp.spawn(_ring_worker, args=(world, port, ret, Bq, Bkv), nprocs=world, join=True)
    q, k, v = _full_inputs()
    full = int8_ref.sage_forward(q, k, v, Bq, Bkv)
    O = torch.cat([ret[r][0] for r in range(world)], dim=2)
    # every rank smoothed K with the same GLOBAL mean (all-reduce of token sums)
    assert torch.equal(ret[0][2], ret[1][2])
    assert (ret[0][2].float() - full[2].float()).abs().max() < 1e-3
    # k-tile visiting order differs per rank -> tolerance, not bitwise
    assert (O.float() - full[0].float()).abs().max() < 6e-3
    lse = torch.cat([ret[r][1].view(1, 2, -1) for r in range(world)], dim=2).reshape(-1)
    ref_lse = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, full[2]), v, Bq, Bkv, return_lse32=True)[10]
    assert (lse - ref_lse).abs().max() < 3e-2


def _shard_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import int8_ref
    from quantizedattention_b200.parallel import shard_batch_heads
    q, k, v = _full_inputs()
    q, k, v = [t.repeat(2, 1, 1, 1) for t in (q, k, v)]             # B*H = 4
    ql, kl, vl = [shard_batch_heads(t, rank, world).contiguous() for t in (q, k, v)]
    out = int8_ref.sage_forward(ql, kl, vl, 128, 128)[0]
    ret[rank] = out
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                        # only timings are ever reduced (bench.py)
    assert t.item() == world
    dist.destroy_process_group()


def test_batch_head_sharding_world2_needs_no_collective():
    from oracle import int8_ref
    world, port = 2, 30100 + (os.getpid() % 500)
    ret = mp.Manager().dict()
    mp.spawn(_shard_worker, args=(world, port, ret), nprocs=world, join=True)
    q, k, v = _full_inputs()
    q, k, v = [t.repeat(2, 1, 1, 1) for t in (q, k, v)]
    full = int8_ref.sage_forward(q, k, v, 128, 128)[0].reshape(4, 512, 64)
    got = torch.cat([ret[r].reshape(2, 512, 64) for r in range(world)], dim=0)
    assert torch.equal(got, full)          # heads are independent: sharding is bit-exact


def test_shard_validation():
    from quantizedattention_b200.parallel import shard_batch_heads
    with pytest.raises(ValueError):
        shard_batch_heads(torch.zeros(1, 3, 128, 64), 0, 2)


def _zigzag_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from quantizedattention_b200.parallel import ring_int8_attention_fwd_causal, zigzag_chunks
    q, k, v = _full_inputs()
    Sc = q.shape[2] // (2 * world)
    a, b = zigzag_chunks(rank, world)
    take = lambda t: torch.cat([t[:, :, a * Sc:(a + 1) * Sc], t[:, :, b * Sc:(b + 1) * Sc]], dim=2).contiguous()
    O, lse, km = ring_int8_attention_fwd_causal(take(q), take(k), take(v), kernels=OracleRingKernels())
    ret[rank] = (O, lse, km)
    dist.barrier()
    dist.destroy_process_group()


def test_causal_zigzag_ring_world2_matches_single_device():
    """causal ring with zig-zag sharding (rank r owns chunks r and 2g-1-r) against the single-device causal
    oracle on the whole sequence (strict mask, global row 0 = uniform average over all keys)."""
    from oracle import int8_ref
    world, port = 2, 30700 + (os.getpid() % 500)
    ret = mp.Manager().dict()
    mp.spawn(_zigzag_worker, args=(world, port, ret), nprocs=world, join=True)
    q, k, v = _full_inputs()
    B, H, S, D = q.shape
    Sc = S // (2 * world)
    full = int8_ref.sage_forward(q, k, v, 128, 128, causal=True)
    ref_lse = int8_ref.int8_fwd(q, int8_ref.smooth_k(k, full[2]), v, 128, 128, return_lse32=True, causal=True)[10].view(B * H, S)
    O = torch.empty_like(full[0])
    lse = torch.empty(B * H, S)
    for r in range(world):
        for half, c in enumerate((r, 2 * world - 1 - r)):
            O[:, :, c * Sc:(c + 1) * Sc] = ret[r][0][:, :, half * Sc:(half + 1) * Sc]
            lse[:, c * Sc:(c + 1) * Sc] = ret[r][1][:, half * Sc:(half + 1) * Sc]
    assert torch.equal(ret[0][2], ret[1][2])
    assert (O.float() - full[0].float()).abs().max() < 6e-3           # per-chunk state continuation: tolerance, not bitwise
    assert (lse - ref_lse).abs().max() < 3e-2
