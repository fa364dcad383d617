"""Per-tile timeline of one CTA of the int8 forward (SM clock stamps recorded by the kernel's roles in the development
library libqattn_dev.so).  Two-stage kernel (nsplit 0 / 2) slots:
  exp warp 0:   0 tile start | 1 logits ready (lg_full) | 2 logits loaded | 5 P handed over
  logit warp 8: 12 tile start | 13 S ready (s_full) | 14 logits + row parameters published
  drain warp:   6 before o_full wait | 7 Opart ready | 8 drained
  MMA thread:   9 K landed, QK issued | 10 V landed and Opart free | 11 P ready, PV issued"""
import os
import sys

os.environ["QA_DEV_LIB"] = "1"      # every call goes through libqattn_dev.so (kernels with timeline hooks)

import torch  # noqa: E402

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib, ops  # noqa: E402

BH, S, D = 148, 8192, 128
torch.manual_seed(0)
q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
L = _lib.lib()
for ns in [int(x) for x in os.environ.get("QA_VARIANTS", "0").split(",")]:
    buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
    ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, nsplit=ns)
    L.qa_debug_set_int8_fwd_timeline(_lib.ptr(buf))
    ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, nsplit=ns)
    torch.cuda.synchronize()
    L.qa_debug_set_int8_fwd_timeline(None)
    t = buf.view(64, 16).cpu()
    t0 = int(t[8, 12])
    print(f"nsplit={ns}: stamps relative to tile 8 logit-warp start (cycles)")
    cols = [(12, "lg_start"), (13, "S_ready"), (14, "lg_done"), (0, "ex_start"), (1, "lg_rdy"), (2, "lg_ld"), (5, "P_done"),
            (10, "V&O_ok"), (11, "PV_iss"), (6, "dr_wait"), (7, "O_ready"), (8, "drained"), (9, "QK_iss")]
    print("tile " + " ".join(f"{n:>8s}" for _, n in cols))
    for j in range(8, 22):
        print(f"{j:4d} " + " ".join(f"{int(t[j, s]) - t0:8d}" for s, _ in cols))
    per_tile = (int(t[40, 5]) - int(t[8, 5])) / 32
    print("cycles per tile (P_done to P_done):", per_tile)
