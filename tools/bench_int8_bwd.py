"""Development: int8 backward kernels at the cfg3 tile shape (B*H heads, S = 8192, D = 128), kernel time from CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402


def run(BH=148, S=8192, D=128, kernel="ws"):
    torch.manual_seed(0)
    q, k, v, dO = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(4)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    doi, sdo = ops.quant_block(dO, 128)
    O, lse16, lse32 = ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D)
    delta = ops.bwd_delta(dO.view(-1, D), O)
    f = lambda: ops.int8_bwd_prequant(qi, ki, vi, doi, sq, sk, sv, sdo, lse32, delta, None, BH, S, D, kernel=kernel)
    for _ in range(3):
        out = f()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    chk = [float(t.float().abs().sum()) for t in out[:3]]
    return {"BH": BH, "S": S, "kernel": kernel, "ms_call": ms, "TOPS_call": 10.0 * BH * S * S * D / ms / 1e9, "checksums": chk}


if __name__ == "__main__":
    print(json.dumps([run(kernel="ws"), run(BH=256, kernel="ws")], indent=1))
