"""Per-tile timeline of one CTA of the int8 forward (SM clock stamps recorded by the kernel's roles).
slots: 0 softmax: before s_full wait | 1 S ready | 2 pass-1 done | 3, 4 scales handed over, pass 2 starts | 5 P handed over
       6 correction: before o_full wait | 7 Opart ready | 8 drained     9 MMA: K landed | 10 QK issued | 11 PV issued"""
import json
import os
os.environ["QA_DEV_LIB"] = "1"      # every call goes through libqattn_dev.so (kernels with timeline hooks)
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib, ops  # noqa: E402

BH, S, D = 148, 8192, 128
torch.manual_seed(0)
q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
L = _lib.lib()
for ns in (2, 1):
    buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
    ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, nsplit=ns)
    L.qa_debug_set_int8_fwd_timeline(_lib.ptr(buf))
    ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, nsplit=ns)
    torch.cuda.synchronize()
    L.qa_debug_set_int8_fwd_timeline(None)
    t = buf.view(64, 16).cpu()
    t0 = int(t[8, 0])
    print(f"nsplit={ns}: stamps relative to tile 8 softmax start (cycles)")
    names = ["sm_wait", "S_ready", "pass1", "pre_pe", "pe_ok", "P_done", "c_wait", "O_ready", "drained", "K_ok", "QK_iss", "PV_iss"]
    print("tile " + " ".join(f"{n:>8s}" for n in names))
    for j in range(8, 20):
        print(f"{j:4d} " + " ".join(f"{int(t[j, s]) - t0:8d}" for s in range(12)))
    per_tile = (int(t[40, 5]) - int(t[8, 5])) / 32
    print("cycles per tile (P_done to P_done):", per_tile)
