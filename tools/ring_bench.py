"""Multi-GPU runs (launch with torchrun, one rank per GPU):
  ring   : BASELINE configs[4] -- int8 long-context forward B=1 H=32 S=131072 D=128, sequence-sharded ring KV
  check  : small ring problem compared against the single-device kernel (parity of the NCCL ring on real GPUs)
  check_causal : the zig-zag causal ring (rank r owns chunks r and 2g-1-r) against the single-device causal kernel
Prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantizedattention_b200 import attention_int8 as A  # noqa: E402
from quantizedattention_b200.parallel import ring_int8_attention_fwd  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:                                        # NCCL's send/recv kernels must not queue behind the attention CTAs that fill every SM
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    except Exception:  # noqa: BLE001
        dist.init_process_group("nccl", device_id=dev)
    if mode == "check_causal":
        from quantizedattention_b200.parallel import ring_int8_attention_fwd_causal, zigzag_chunks
        B, H, S, D = 1, 4, 4096, 128
        g = torch.Generator().manual_seed(1006)
        q, k, v = [torch.randn(B, H, S, D, generator=g).to(torch.float16) for _ in range(3)]
        Sc = S // (2 * world)
        a, b = zigzag_chunks(rank, world)
        take = lambda t: torch.cat([t[:, :, a * Sc:(a + 1) * Sc], t[:, :, b * Sc:(b + 1) * Sc]], dim=2).contiguous().to(dev)
        O, lse, _ = ring_int8_attention_fwd_causal(take(q), take(k), take(v))
        ref = A.sage_attention_3_int8(q.to(dev), k.to(dev), v.to(dev), causal=True)
        mine = torch.cat([ref[:, :, a * Sc:(a + 1) * Sc], ref[:, :, b * Sc:(b + 1) * Sc]], dim=2)
        err = (O.float() - mine.float()).abs().max()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"mode": mode, "n_gpus": world, "S": S, "max_abs_vs_single_device": err.item(), "ok": bool(err.item() < 6e-3)}))
        dist.destroy_process_group()
        return
    if mode == "check":
        B, H, S, D = 1, 4, 4096, 128
    else:
        B, H, S, D = 1, 32, 131072, 128
    Sl = S // world
    g = torch.Generator().manual_seed(1005)
    if mode == "check":
        q, k, v = [torch.randn(B, H, S, D, generator=g).to(torch.float16) for _ in range(3)]
        ql, kl, vl = [t[:, :, rank * Sl:(rank + 1) * Sl].contiguous().to(dev) for t in (q, k, v)]
    else:
        g = torch.Generator(device=dev).manual_seed(1005 + rank)
        ql, kl, vl = [torch.randn(B, H, Sl, D, generator=g, device=dev, dtype=torch.float16) for _ in range(3)]
    for _ in range(2):
        out = ring_int8_attention_fwd(ql, kl, vl)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 3
    e0.record()
    for _ in range(iters):
        out = ring_int8_attention_fwd(ql, kl, vl)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"mode": mode, "n_gpus": world, "B": B, "H": H, "S": S, "D": D, "ms": t.item(),
           "TOPS_total": 4.0 * B * H * S * S * D / (t.item() * 1e-3) / 1e12}
    if mode == "check":
        ref = A.SageAttention3_Int8_autograd_function.forward(q.to(dev), k.to(dev), v.to(dev))
        err = (out[0].float() - ref[0][:, :, rank * Sl:(rank + 1) * Sl].float()).abs().max()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        res["max_abs_vs_single_device"] = err.item()
        res["ok"] = bool(err.item() < 6e-3)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
