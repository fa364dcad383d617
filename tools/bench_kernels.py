"""Kernel-level timing (CUDA events) for development.  Usage: python tools/bench_kernels.py int8_fwd [BH S D]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def int8_fwd(BH=256, S=8192, D=128):
    torch.manual_seed(0)
    q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
    res = {}
    med, best = timeit(lambda: ops.quant_block(q, 128))
    res["quant_q_GBs"] = 3 * q.numel() / med / 1e6
    q_i8, sq = ops.quant_block(q, 128); k_i8, sk = ops.quant_block(k, 128); v_i8, sv = ops.quant_block(v, 128)
    km = ops.k_mean(k.view(1, BH, S, D))
    med, _ = timeit(lambda: ops.k_mean(k.view(1, BH, S, D)))
    res["k_mean_GBs"] = 2 * k.numel() / med / 1e6
    med, _ = timeit(lambda: ops.quant_block(k, 128, mean=km, rows_per_head=S))
    res["quant_k_smooth_GBs"] = 3 * k.numel() / med / 1e6
    for ns in (1, 2):
        med, best = timeit(lambda: ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, S, S, D, 128, 128, nsplit=ns))
        res[f"int8_fwd_nsplit{ns}"] = {"ms_med": med, "ms_best": best, "TOPS_med": 4 * BH * S * S * D / med / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[2:]]
    globals()[sys.argv[1]](*args)
