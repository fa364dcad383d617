"""Development: NVFP4 forward kernel and its quantisation pre-passes at B*H = 64, S = 8192, D = 128 (CUDA events)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import attention_fp4 as F  # noqa: E402
from quantizedattention_b200 import ops  # noqa: E402


def run(B=2, H=32, S=8192, D=128, variant=0):
    torch.manual_seed(0)
    q, k, v = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
    o = F.quantise_fp4(q, k, v)
    for _ in range(3):
        F.fp4_fwd_prequant(o, variant)
    ops.TIMING = []
    for _ in range(10):
        F.fp4_fwd_prequant(o, variant)
    for _ in range(5):
        F.quantise_fp4(q, k, v)
    torch.cuda.synchronize()
    med = lambda name: sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == name)[len([1 for n, _, _ in ops.TIMING if n == name]) // 2]
    ms, mq, mv = med("fp4_fwd"), med("fp4_quant_rows"), med("fp4_quant_vt")
    ops.TIMING = None
    nbytes = B * H * S * D * 2.5625
    return {"BH": B * H, "S": S, "variant": variant, "ms_kernel": ms, "TFLOPS": 4.0 * B * H * S * S * D / ms / 1e9, "quant_rows_ms": mq,
            "quant_rows_GBs": (nbytes + B * H * S * D * 2) / mq / 1e6, "quant_vt_ms": mv, "quant_vt_GBs": nbytes / mv / 1e6}


if __name__ == "__main__":
    print(json.dumps([run(), run(variant=1), run(8, 32, 8192, 128)], indent=1))
