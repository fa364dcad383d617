"""Timeline of the host-staged pipeline (per-chunk H2D / compute / D2H spans, ms since the first event)."""
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200.host_pipeline import HostStagedSageAttention  # noqa: E402

hc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, H, S, D = 8, 32, 8192, 128
g = torch.Generator().manual_seed(0)
host = [torch.randn(B, H, S, D, generator=g).half().pin_memory() for _ in range(4)]
out = [torch.empty(B, H, S, D, dtype=torch.float16).pin_memory() for _ in range(4)]
pipe = HostStagedSageAttention(heads_per_chunk=hc, slots=slots)
for _ in range(2):
    pipe(*host, out=out)
torch.cuda.synchronize()
import time
for rep in range(3):                                    # untraced: host time to enqueue one step vs device time
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    pipe(*host, out=out)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("hc", hc, "slots", slots, "device ms", round(a.elapsed_time(b), 2), "host enqueue ms", round((t1 - t0) * 1e3, 2))
pipe.trace = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
pipe(*host, out=out)
b.record()
torch.cuda.synchronize()
print("hc", hc, "slots", slots, "total ms", round(a.elapsed_time(b), 2))
for st, c, s, e in pipe.trace:
    print(f"{st:8s} {c:3d}  {a.elapsed_time(s):8.2f} -> {a.elapsed_time(e):8.2f}  ({s.elapsed_time(e):6.2f} ms)")
