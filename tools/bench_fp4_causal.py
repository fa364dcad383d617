"""Development: NVFP4 forward, causal vs dense, B*H = 64, S = 8192, D = 128."""
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import attention_fp4 as F, ops  # noqa: E402

for B in (2, 8):
  q, k, v = [torch.randn(B, 32, 8192, 128, device="cuda", dtype=torch.float16) for _ in range(3)]
  o = F.quantise_fp4(q, k, v)
  for c in (False, True):
    for _ in range(3):
        F.fp4_fwd_prequant(o, causal=c)
    ops.TIMING = []
    for _ in range(7):
        F.fp4_fwd_prequant(o, causal=c)
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for n, a, b in ops.TIMING if n == "fp4_fwd")[3]
    ops.TIMING = None
    print("B*H", B * 32, "causal", c, "ms", round(t, 3), "TFLOPS (causal = half of dense)", round((0.5 if c else 1) * 4 * B * 32 * 8192 * 8192 * 128 / t / 1e9, 1))
