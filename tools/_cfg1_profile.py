"""Development: host-side cost of one sage_attention_3_int8 call at configs[0] (B=1 H=8 S=1024 D=64)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import attention_int8 as A  # noqa: E402

q, k, v = [torch.randn(1, 8, 1024, 64, device="cuda", dtype=torch.float16) for _ in range(3)]
with torch.no_grad():
    for _ in range(20):
        A.sage_attention_3_int8(q, k, v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        A.sage_attention_3_int8(q, k, v)
    torch.cuda.synchronize()
    print("ms per call", (time.perf_counter() - t0) / 200 * 1e3)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        A.sage_attention_3_int8(q, k, v)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
