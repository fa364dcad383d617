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
    for ns in (0, 1, 2):
        med, best = timeit(lambda: ops.int8_fwd_prequant(q_i8, k_i8, v_i8, sq, sk, sv, BH, S, S, D, 128, 128, nsplit=ns))
        res[f"int8_fwd_nsplit{ns}"] = {"ms_med": med, "ms_best": best, "TOPS_med": 4 * BH * S * S * D / med / 1e9}
    print(json.dumps(res, indent=1))


def bf16(B=4, H=16, S=4096, D=128, causal=1):
    torch.manual_seed(0)
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B, H, S, D, device="cuda")
    res = {}
    f = 0.5 if causal else 1.0
    for ns in (1, 2, 3):
        med, best = timeit(lambda: ops.bf16_fwd(q, k, v, bool(causal), nsplit=ns))
        res[f"bf16_fwd_nsplit{ns}"] = {"ms": med, "TFLOPS": f * 4 * B * H * S * S * D / med / 1e9}
    O, lse = ops.bf16_fwd(q, k, v, bool(causal))
    ops.TIMING = []
    med, best = timeit(lambda: ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO))
    kt = [a.elapsed_time(b) for n, a, b in ops.TIMING if n == "bf16_bwd"]
    ops.TIMING = None
    kms = sorted(kt)[len(kt) // 2]
    res["bf16_bwd"] = {"ms_total": med, "ms_kernel": kms, "TFLOPS_kernel": f * 10 * B * H * S * S * D / kms / 1e9}
    print(json.dumps(res, indent=1))


def jvp(B=16, H=16, S=4096, D=64):
    torch.manual_seed(0)
    t = [torch.randn(B, H, S, D, device="cuda") for _ in range(6)]
    res = {}
    for ns in (1, 2):
        ops.TIMING = []
        med, best = timeit(lambda: ops.jvp_fwd(*t, nsplit=ns))
        kt = sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == "jvp_fwd")
        ops.TIMING = None
        kms = kt[len(kt) // 2]
        res[f"jvp_nsplit{ns}"] = {"ms_total": med, "ms_kernel": kms, "TFLOPS_kernel": 12 * B * H * S * S * D / kms / 1e9}
    print(json.dumps(res, indent=1))


def int8_causal(BH=256, S=8192, D=128):
    """Causal int8 forward + backward kernels; ops counted as half of the dense convention."""
    torch.manual_seed(0)
    q, k, v, dO = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(4)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    doi, sdo = ops.quant_block(dO, 128)
    res = {}
    for causal in (False, True):
        f = 0.5 if causal else 1.0
        fwd = lambda: ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, causal=causal)
        med, _ = timeit(fwd)
        O, lse16, lse32 = fwd()
        delta = ops.bwd_delta(dO.view(-1, D), O)
        ops.TIMING = []
        timeit(lambda: ops.int8_bwd_prequant(qi, ki, vi, doi, sq, sk, sv, sdo, lse32, delta, None, BH, S, D, causal=causal))
        kt = sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == "int8_bwd")
        ops.TIMING = None
        kms = kt[len(kt) // 2]
        res["causal" if causal else "dense"] = {"fwd_ms": med, "fwd_TOPS": f * 4 * BH * S * S * D / med / 1e9,
                                                "bwd_ms": kms, "bwd_TOPS": f * 10 * BH * S * S * D / kms / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[2:]]
    globals()[sys.argv[1]](*args)
