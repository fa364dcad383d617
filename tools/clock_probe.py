"""Sample SM clock / power / throttle reasons while one kernel runs back to back.  Usage: python tools/clock_probe.py bf16_fwd|int8_fwd|int8_bwd"""
import json
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "bf16_fwd"
torch.manual_seed(0)
if which == "bf16_fwd":
    B, H, S, D = 8, 32, 8192, 128
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.bf16_fwd(q, k, v, False)
    flops = 4 * B * H * S * S * D
else:
    BH, S, D = 256, 8192, 128
    q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    fn = lambda: ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D)
    flops = 4 * BH * S * S * D
samples, stop = [], False


def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,temperature.gpu",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.05)


for _ in range(3):
    fn()
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 150
a.record()
for _ in range(n):
    fn()
b.record(); torch.cuda.synchronize()
stop = True; th.join()
ms = a.elapsed_time(b) / n
print(json.dumps({"kernel": which, "ms": ms, "TFLOPS": flops / ms / 1e9, "samples": samples[::max(1, len(samples) // 12)]}, indent=1))
