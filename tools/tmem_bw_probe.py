"""Measure TMEM->register read bandwidth per SM (bytes/clk) for 4..32 warps per CTA, one CTA per SM."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib  # noqa: E402

L = _lib.dev_lib()
sink = torch.zeros(2048, dtype=torch.int32, device="cuda")
props = torch.cuda.get_device_properties(0)
sms = props.multi_processor_count
out = {}
SHAPES = {0: "32x32b.x32", 1: "16x256b.x8", 2: "16x128b.x16", 3: "16x64b.x32"}
for shape, depth, threads in [(s, d, t) for s in SHAPES for d in (1, 2) for t in (128, 256, 512, 1024) if not (d == 2 and t > 512)]:
    iters = 5000
    L.qa_probe_tmem_bw_ex(_lib.ptr(sink), sms, threads, 100, shape, depth, _lib.cur_stream())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(L.qa_probe_tmem_bw_ex(_lib.ptr(sink), sms, threads, iters, shape, depth, _lib.cur_stream()), "probe", L)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    bytes_per_sm = (threads // 32) * iters * 4 * 4096
    clk_hz = 1.965e9
    out[f"{SHAPES[shape]}_depth{depth}_{threads}thr"] = {"ms": round(ms, 3), "bytes_per_clk_per_sm_at_1965MHz": round(bytes_per_sm / (ms * 1e-3) / clk_hz, 1)}
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/tmem_bw.json", "w"), indent=1)
