"""Development: bf16-path backward kernels at configs[1] (B=4 H=16 S=4096 D=128 causal) and non-causal S=8192."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402


def run(B, H, S, D, causal, variant):
    torch.manual_seed(0)
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B, H, S, D, device="cuda")
    O, lse = ops.bf16_fwd(q, k, v, bool(causal))
    for _ in range(3):
        ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO, variant=variant)
    ops.TIMING = []
    for _ in range(10):
        ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO, variant=variant)
    torch.cuda.synchronize()
    kt = sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == "bf16_bwd")
    ops.TIMING = None
    ms = kt[len(kt) // 2]
    f = 0.5 if causal else 1.0
    return {"B": B, "H": H, "S": S, "D": D, "causal": causal, "variant": variant, "ms_kernel": ms,
            "TFLOPS": f * 10 * B * H * S * S * D / ms / 1e9}


if __name__ == "__main__":
    out = []
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":       # 37 heads: whole waves of 148 CTAs; time per wave = overhead + tiles * t_tile
        for S in (1024, 2048, 4096, 8192):
            r = run(1, 37, S, 128, 0, 0)
            waves = 37 * (S // 128) / 148
            r["us_per_wave"] = r["ms_kernel"] * 1e3 / waves
            r["tiles_per_cta"] = S // 128
            out.append(r)
        print(json.dumps(out, indent=1))
        sys.exit(0)
    for variant in (0, 1):
        out.append(run(4, 16, 4096, 128, 1, variant))
        out.append(run(1, 32, 8192, 128, 0, variant))
    print(json.dumps(out, indent=1))
