"""Development: kernel times of the main kernels in one go (A/B runs of build variants)."""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from quantizedattention_b200 import ops  # noqa: E402
import bench_bf16_bwd  # noqa: E402
import bench_bf16_fwd  # noqa: E402
import bench_int8_bwd  # noqa: E402


def int8_fwd(BH=148, S=8192, D=128):
    q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    f = lambda: ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D)
    for _ in range(3):
        f()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    return {"ms_call": ms, "TOPS": 4.0 * BH * S * S * D / ms / 1e9}


if __name__ == "__main__":
    r = {"int8_fwd_TOPS": round(int8_fwd()["TOPS"], 1), "int8_bwd_TOPS": round(bench_int8_bwd.run(BH=148)["TOPS_call"], 1),
         "bf16_bwd_cfg2_TFLOPS": round(bench_bf16_bwd.run(4, 16, 4096, 128, 1, 0)["TFLOPS"], 1),
         "bf16_bwd_S8k_TFLOPS": round(bench_bf16_bwd.run(1, 32, 8192, 128, 0, 0)["TFLOPS"], 1),
         "bf16_fwd_cfg2_TFLOPS": round(bench_bf16_fwd.run(4, 16, 4096, 128, 1)["TFLOPS"], 1),
         "bf16_fwd_S8k_TFLOPS": round(bench_bf16_fwd.run(1, 64, 8192, 128, 0)["TFLOPS"], 1)}
    print("TOPS " + json.dumps(r))
