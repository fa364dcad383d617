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

    def attend_causal_diag(self, q_i8, sq, kv, BH, S, D):
        """The diagonal chunk of a causal ring: strict mask key < query inside the chunk, fresh state in, state out."""
        raise NotImplementedError

    def v_token_sum(self, v_i8, sv, BH, S, D, Bkv):               # -> fp32 [BH, D]: sum over tokens of the de-quantised V
        return (v_i8.view(BH, S // Bkv, Bkv, D).float() * sv.view(BH, S // Bkv, 1, 1).float()).sum(dim=(1, 2))


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

    def attend_causal_diag(self, q_i8, sq, kv, BH, S, D):
        from . import ops
        k_i8, v_i8, sk, sv = kv
        return ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, S, S, D, 128, 128, ring_state=True, causal=True)


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


def zigzag_chunks(rank: int, world: int):
    """Chunk ids (of 2 * world equal chunks of the sequence) a rank owns under zig-zag sharding: (rank, 2*world-1-rank).
    Every rank then has the same amount of causal work."""
    return rank, 2 * world - 1 - rank


def ring_int8_attention_fwd_causal(q, k, v, group=None, kernels: RingInt8Kernels | None = None):
    """Causal int8 attention forward, sequence-sharded with zig-zag chunks (causal ring for long-context
    training).  The sequence is cut into 2g chunks; rank r owns chunks r and 2g-1-r, and passes q, k, v as
    fp16 [B,H,2*Sc,D] = [chunk r | chunk 2g-1-r].  The mask is the strict one of the reference's baseline (key < query,
    attention_int8.py:465-473; global row 0 = uniform average over all keys of the de-quantised V, LEDGER B-1).
    Bq = Bkv = 128.  Returns (O fp16 [B,H,2*Sc,D] in the same chunk order, lse32 fp32 [B*H, 2*Sc], k_mean).

    Per ring step a rank holds the K/V chunks (c, 2g-1-c) of rank c = r - step: for c < r both of its query chunks attend
    chunk c in full; for c > r only its late query chunk attends both; its own pair is the diagonal case (strict mask
    inside the two diagonal chunks).  Every (query chunk, K/V chunk) product continues the same online-softmax state in
    the kernel, so nothing is merged on the host; chunk pairs above the diagonal are never launched."""
    kernels = kernels or CudaRingKernels()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, S2, D = q.shape
    Sc, BH = S2 // 2, B * H
    if S2 % 256:
        raise ValueError("every zig-zag chunk must be a multiple of 128 tokens")
    k_mean = global_k_mean(k, S2 * world, kernels.token_sum, group)
    halves = lambda t: (t[:, :, :Sc].contiguous(), t[:, :, Sc:].contiguous())
    qa, qb = [kernels.quant(x, 128) for x in halves(q)]
    (ka, ska), (kb, skb) = [kernels.quant(x, 128, mean=k_mean, rows_per_head=Sc) for x in halves(k)]
    (va, sva), (vb, svb) = [kernels.quant(x, 128) for x in halves(v)]
    cur = (ka, va, ska, sva, kb, vb, skb, svb)                       # K/V chunk pair currently held (own pair first)
    # global row 0 (LEDGER B-1): sum over ALL tokens of the de-quantised V
    vsum = kernels.v_token_sum(va, sva, BH, Sc, D, 128) + kernels.v_token_sum(vb, svb, BH, Sc, D, 128)
    if dist.is_initialized() and world > 1:
        dist.all_reduce(vsum, op=dist.ReduceOp.SUM, group=group)
    use_cuda = q.is_cuda
    comm_stream = torch.cuda.Stream(device=q.device) if (use_cuda and world > 1) else None
    st_a = st_b = None
    out_a = out_b = None
    for step in range(world):
        src = (rank - step) % world                                 # owner of the K/V pair held in this step
        nxt, reqs = None, []
        if step + 1 < world:
            nxt = tuple(torch.empty_like(t) for t in cur)
            send_to, recv_from = (rank + 1) % world, (rank - 1) % world
            g_rank = (lambda r_: dist.get_global_rank(group, r_)) if group else (lambda r_: r_)
            ops_ = [dist.P2POp(dist.isend, t, g_rank(send_to), group) for t in cur]
            ops_ += [dist.P2POp(dist.irecv, t, g_rank(recv_from), group) for t in nxt]
            if comm_stream is not None:
                comm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm_stream):
                    reqs = dist.batch_isend_irecv(ops_)
            else:
                reqs = dist.batch_isend_irecv(ops_)
        kv_lo, kv_hi = cur[:4], cur[4:]                             # chunks src and 2g-1-src
        # which call is the LAST one of each query chunk: chunk a ends at step = rank (src = 0), chunk b at the last step
        last_a = (step == rank)
        last_b = (step == world - 1)
        if step == 0:                                               # own pair: the two diagonal chunks and (b, a) in full
            st_a = kernels.attend_causal_diag(qa[0], qa[1], kv_lo, BH, Sc, D)
            if last_a:                                              # rank 0: chunk 0 sees nothing else; normalise its state
                out_a = _finish_state(st_a)
            st_b = kernels.attend_causal_diag(qb[0], qb[1], kv_hi, BH, Sc, D)
            res = kernels.attend(qb[0], qb[1], kv_lo, st_b, BH, Sc, Sc, D, 128, 128, last_b)
            st_b, out_b = (None, res) if last_b else (res, None)
        elif src < rank:                                            # an earlier rank's pair: both query chunks see chunk src
            res = kernels.attend(qa[0], qa[1], kv_lo, st_a, BH, Sc, Sc, D, 128, 128, last_a)
            st_a, out_a = (None, res) if last_a else (res, out_a)
            res = kernels.attend(qb[0], qb[1], kv_lo, st_b, BH, Sc, Sc, D, 128, 128, last_b)
            st_b, out_b = (None, res) if last_b else (res, None)
        else:                                                       # a later rank's pair: only the late query chunk, both chunks
            st_b = kernels.attend(qb[0], qb[1], kv_lo, st_b, BH, Sc, Sc, D, 128, 128, False)
            res = kernels.attend(qb[0], qb[1], kv_hi, st_b, BH, Sc, Sc, D, 128, 128, last_b)
            st_b, out_b = (None, res) if last_b else (res, None)
        if comm_stream is not None:
            with torch.cuda.stream(comm_stream):
                for r_ in reqs:
                    r_.wait()
            torch.cuda.current_stream().wait_stream(comm_stream)
        else:
            for r_ in reqs:
                r_.wait()
        if nxt is not None:
            cur = nxt
    Oa, _, lsa = out_a
    Ob, _, lsb = out_b
    O = torch.cat([Oa.view(B, H, Sc, D), Ob.view(B, H, Sc, D)], dim=2)
    lse = torch.cat([lsa.view(BH, Sc), lsb.view(BH, Sc)], dim=1)
    if rank == 0:                                                   # global row 0 sees no key: uniform average over all keys
        import math
        S_tot = S2 * world
        O[:, :, 0] = (vsum / S_tot).view(B, H, D).to(O.dtype)
        lse[:, 0] = -128.0 + math.log2(S_tot)
    return O, lse, k_mean


def _finish_state(state):
    """(O_acc, m, l) -> (O fp16, lse16, lse32): the normalisation the kernel applies on a `last` call, for the one chunk
    whose only product is a diagonal one (chunk 0 of rank 0)."""
    o_acc, m, l = state
    lse32 = m + torch.log2(l)
    return (o_acc / l[:, None]).to(torch.float16), lse32.to(torch.float16), lse32
