"""NOTE: run with QA_INT8_BWD_WS=0: the stamps live in the 8-warp kernel; the default warp-specialised kernel carries
none (they cost 3-4 % through register pressure).

Per-q-tile timeline of one CTA of the int8 backward (SM clock stamps of the leader warp 0 and of warp 5).
slots: 0 loop top | 1 S/dP ready | 2 pass 1 done | 3 dV/dK partial ready | 4 dV/dK drained | 5 barrier 1 passed |
       6 dQ issued (leader) | 7 pass 2 done | 8 dQ partial ready | 9 dQ drained | 10 barrier 2 passed |
       11 S/dP(t+1) issued (leader) | 12 dV/dK issued (leader) = end of iteration"""
import os
import sys

os.environ["QA_DEV_LIB"] = "1"      # every call goes through libqattn_dev.so (kernels with timeline hooks)

import torch  # noqa: E402

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib, ops  # noqa: E402

BH, S, D = 148, 8192, 128
torch.manual_seed(0)
q, k, v, dO = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(4)]
qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
doi, sdo = ops.quant_block(dO, 128)
O, lse16, lse32 = ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D)
delta = ops.bwd_delta(dO.view(-1, D), O)
L = _lib.lib()
args = (qi, ki, vi, doi, sq, sk, sv, sdo, lse32, delta, None, BH, S, D)
ops.int8_bwd_prequant(*args)
buf = torch.zeros(64 * 2 * 16, dtype=torch.int64, device="cuda")
L.qa_debug_set_int8_bwd_timeline(_lib.ptr(buf))
ops.int8_bwd_prequant(*args)
torch.cuda.synchronize()
L.qa_debug_set_int8_bwd_timeline(None)
t = buf.view(64, 2, 16).cpu()
if os.environ.get("QA_INT8_BWD_WS", "1") != "0":        # warp-specialised kernel: quantise warp 1 (slots 0-5), leader warp 8 (6-15)
    qn = ["top", "S_rdy", "pass1", "amax_bar", "P_free", "pass2"]
    dn = ["top", "dVK_rdy", "dVK_drn", "bar1", "dQ+S_iss", "dQ_rdy", "dQ_drn", "bar2", "pds_ok", "iss_end"]
    print("quantise warp 1 - stamps relative to the tile's loop top (cycles); last column = iteration length")
    print("tile " + " ".join(f"{n:>8s}" for n in qn) + "     iter")
    for j in range(20, 28):
        t0 = int(t[j, 0, 0])
        print(f"{j:4d} " + " ".join(f"{int(t[j, 0, s]) - t0:8d}" for s in range(6)) + f" {int(t[j + 1, 0, 0]) - t0:8d}")
    print("pass-2 end of quantise warps 0..7 (relative to warp 0 loop top)")
    for j in range(20, 28):
        t0 = int(t[j, 0, 0])
        print(f"{j:4d} " + " ".join(f"{int(t[j, 0, 6 + w]) - t0:8d}" for w in range(8)))
    print("leader warp 8 (drain role) - relative to the quantise warp's loop top of the same tile")
    print("tile " + " ".join(f"{n:>8s}" for n in dn) + "     iter")
    for j in range(20, 28):
        t0 = int(t[j, 0, 0])
        print(f"{j:4d} " + " ".join(f"{int(t[j, 1, s]) - t0:8d}" for s in range(6, 16)) + f" {int(t[j + 1, 1, 6]) - int(t[j, 1, 6]):8d}")
    print("leader: before pds_full wait / after wait / after fence / issue end (relative to the quantise loop top)")
    for j in range(20, 28):
        t0 = int(t[j, 0, 0])
        print(f"{j:4d} " + " ".join(f"{int(t[j, 1, s]) - t0:8d}" for s in (0, 1, 14, 15)))
    print("cycles per tile:", (int(t[52, 0, 0]) - int(t[20, 0, 0])) / 32)
    sys.exit(0)
names = ["top", "S_rdy", "pass1", "dVK_rdy", "dVK_drn", "bar1", "dQ_iss", "pass2", "dQ_rdy", "dQ_drn", "bar2", "SdP_iss", "dVK_iss"]
for w, label in ((0, "leader warp 0"), (1, "warp 5")):
    print(label, "- stamps relative to the tile's loop top (cycles); last column = iteration length")
    print("tile " + " ".join(f"{n:>8s}" for n in names) + "     iter")
    for j in range(20, 28):
        t0 = int(t[j, w, 0])
        row = " ".join(f"{(int(t[j, w, s]) - t0) if int(t[j, w, s]) else 0:8d}" for s in range(13))
        print(f"{j:4d} {row} {int(t[j + 1, w, 0]) - t0:8d}")
print("cycles per tile:", (int(t[52, 0, 0]) - int(t[20, 0, 0])) / 32)
