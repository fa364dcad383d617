"""Development: bf16-path forward at configs[1] (B=4 H=16 S=4096 D=128 causal) and non-causal B*H=64 S=8192."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402


def run(B, H, S, D, causal):
    torch.manual_seed(0)
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.bf16_fwd(q, k, v, bool(causal))
    ops.TIMING = []
    for _ in range(10):
        ops.bf16_fwd(q, k, v, bool(causal))
    torch.cuda.synchronize()
    kt = sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == "bf16_fwd")
    ops.TIMING = None
    ms = kt[len(kt) // 2]
    f = 0.5 if causal else 1.0
    return {"B": B, "H": H, "S": S, "D": D, "causal": causal, "ms_kernel": ms, "TFLOPS": f * 4 * B * H * S * S * D / ms / 1e9}


if __name__ == "__main__":
    print(json.dumps([run(4, 16, 4096, 128, 1), run(1, 64, 8192, 128, 0), run(4, 16, 4096, 128, 0), run(16, 16, 4096, 64, 0)], indent=1))
