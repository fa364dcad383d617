"""Multi-GPU partitioning of the attention paths (SURVEY.md 8 row e).  One process per GPU, torch.distributed (NCCL
over NVLink 5 / NVSwitch on the GPU box; gloo on CPU in the tests).

  * batch x head sharding (configs 2, 3, 4): every (b, h) is an independent attention problem -> `shard_batch_heads`
    slices the B*H axis; there is NO collective on the data path.
  * sequence-sharded ring KV (config 5, non-causal int8 forward): rank r owns the query rows and the K/V rows
    [r*S/g, (r+1)*S/g).  One tiny all-reduce (K token sums, [B,H,1,D] fp32) makes every rank smooth K with the same
    global mean; then g steps: attend the local queries to the K/V shard currently held, continuing the SAME
    online-softmax state (m, l, O) across steps, while the int8 K/V shard + its scales travel to the next rank with
    batched isend/irecv on a side stream (double-buffered, overlapping the kernel).  The attention kernel keeps one CTA
    on every SM, so create the NCCL process group with high-priority streams
    (`ProcessGroupNCCL.Options(is_high_priority_stream=True)`, as bench.py does): NCCL's send/recv CTAs then take the
    next SM an attention CTA leaves instead of queueing behind the whole grid.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch_heads(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """[B,H,S,D] -> this rank's contiguous slice of the flattened B*H axis, shape [1, B*H/world, S, D]."""
    B, H, S, D = t.shape
    n = B * H
    if n % world:
        raise ValueError(f"B*H = {n} is not divisible by the world size {world}")
    per = n // world
    return t.reshape(n, S, D)[rank * per:(rank + 1) * per].unsqueeze(0)


def global_k_mean(k_local: torch.Tensor, S_total: int, token_sum_fn, group=None) -> torch.Tensor:
    """fp16 [B,H,1,D] mean of K over the GLOBAL sequence: all-reduce of the per-shard fp32 token sums."""
    s = token_sum_fn(k_local)                      # fp32 [B,H,1,D]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return (s / float(S_total)).to(torch.float16)


class RingInt8Kernels:
    """The device operations the ring needs; the GPU implementation binds the C-ABI kernels, the CPU tests bind the oracle."""

    def token_sum(self, k):                                  # -> fp32 [B,H,1,D]
        raise NotImplementedError

    def quant(self, x, blk, mean=None, rows_per_head=None):  # -> (int8 [N,D], fp16 scales)
        raise NotImplementedError

    def attend(self, q_i8, sq, kv, state, BH, Sq, Sk, D, Bq, Bkv, last):
        """Attend the local queries to one K/V shard `kv = (k_i8, v_i8, sk, sv)`, continuing `state`
        (None or (O_acc, m, l)).  last=False -> new state; last=True -> (O fp16, lse16, lse32)."""
        raise NotImplementedError


class CudaRingKernels(RingInt8Kernels):
    def token_sum(self, k):
        from . import ops
        return ops.k_token_sum(k)

    def quant(self, x, blk, mean=None, rows_per_head=None):
        from . import ops
        return ops.quant_block(x, blk, mean=mean, rows_per_head=rows_per_head)

    def attend(self, q_i8, sq, kv, state, BH, Sq, Sk, D, Bq, Bkv, last):
        from . import ops
        k_i8, v_i8, sk, sv = kv
        return ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, Sq, Sk, D, Bq, Bkv, ring_state=not last,
                                     state_in=state, want_lse32=True)


def ring_int8_attention_fwd(q, k, v, Bq: int = 128, Bkv: int = 128, group=None, kernels: RingInt8Kernels | None = None,
                            timing: list | None = None):
    """Sequence-sharded int8 attention forward.  q, k, v: this rank's fp16 [B,H,S/g,D] shards (rank order = sequence
    order).  Returns (O fp16 [B,H,S/g,D], lse fp16 [B*H*S/g], lse32, k_mean fp16 [B,H,1,D]).
    timing: pass a list to collect one (step, compute_start, compute_end, comm_start, comm_end) tuple of CUDA events per
    ring step (comm events are None on the last step): bench.py reports the kernel time, the send/recv time and how much
    of the latter the kernel hides."""
    kernels = kernels or CudaRingKernels()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, Sl, D = q.shape
    BH = B * H
    k_mean = global_k_mean(k, Sl * world, kernels.token_sum, group)
    q_i8, sq = kernels.quant(q, Bq)
    k_i8, sk = kernels.quant(k, Bkv, mean=k_mean, rows_per_head=Sl)
    v_i8, sv = kernels.quant(v, Bkv)
    cur = (k_i8, v_i8, sk, sv)
    state = None
    use_cuda = q.is_cuda
    comm_stream = torch.cuda.Stream(device=q.device) if (use_cuda and world > 1) else None
    out = None
    for step in range(world):
        nxt, reqs = None, []
        if step + 1 < world:                                   # pass the shard we hold to rank+1, receive from rank-1
            nxt = tuple(torch.empty_like(t) for t in cur)
            send_to, recv_from = (rank + 1) % world, (rank - 1) % world
            ops_ = [dist.P2POp(dist.isend, t, dist.get_global_rank(group, send_to) if group else send_to, group) for t in cur]
            ops_ += [dist.P2POp(dist.irecv, t, dist.get_global_rank(group, recv_from) if group else recv_from, group) for t in nxt]
            if comm_stream is not None:
                comm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm_stream):
                    if timing is not None:
                        ev_c0 = torch.cuda.Event(enable_timing=True); ev_c0.record()
                    reqs = dist.batch_isend_irecv(ops_)
            else:
                reqs = dist.batch_isend_irecv(ops_)
        last = step == world - 1
        if timing is not None and use_cuda:
            ev_k0 = torch.cuda.Event(enable_timing=True); ev_k0.record()
        res = kernels.attend(q_i8, sq, cur, state, BH, Sl, Sl, D, Bq, Bkv, last)
        if timing is not None and use_cuda:
            ev_k1 = torch.cuda.Event(enable_timing=True); ev_k1.record()
        if last:
            out = res
        else:
            state = res
        if comm_stream is not None:
            ev_c1 = None
            with torch.cuda.stream(comm_stream):               # the side stream waits for the transfers, the kernel does not
                for r in reqs:
                    r.wait()
                if timing is not None and nxt is not None:
                    ev_c1 = torch.cuda.Event(enable_timing=True); ev_c1.record()
            if timing is not None:
                timing.append((step, ev_k0, ev_k1, ev_c0 if ev_c1 is not None else None, ev_c1))
            torch.cuda.current_stream().wait_stream(comm_stream)   # the next step's kernel reads the received shard
        else:
            for r in reqs:
                r.wait()
            if timing is not None and use_cuda:
                timing.append((step, ev_k0, ev_k1, None, None))
        if nxt is not None:
            cur = nxt
    O, lse16, lse32 = out
    return O.view(B, H, Sl, D), lse16, lse32, k_mean
