"""Small single-kernel targets for ncu.  Usage: python tools/prof_target.py bf16_fwd|jvp|int8_fwd|bf16_bwd|int8_bwd|fp4_fwd"""
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import ops  # noqa: E402

which = sys.argv[1]
torch.manual_seed(0)
if which == "bf16_fwd":
    B, H, S, D = 4, 16, 4096, 128
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.bf16_fwd(q, k, v, False)
elif which == "jvp":
    t = [torch.randn(4, 16, 4096, 64, device="cuda") for _ in range(6)]
    for _ in range(3):
        ops.jvp_fwd(*t)
elif which == "int8_fwd":
    import os
    BH, S, D = 37, 8192, 128                                   # 37 heads x 64 q-tiles = 16 waves of 148 CTAs
    q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    for _ in range(3):
        ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, nsplit=int(os.environ.get("QA_NSPLIT", "0")))
elif which == "bf16_bwd":
    B, H, S, D = 4, 16, 4096, 128
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B, H, S, D, device="cuda")
    O, lse = ops.bf16_fwd(q, k, v, True)
    for _ in range(3):
        ops.bf16_bwd(q, k, v, O, lse, True, dO)
elif which == "int8_bwd":
    BH, S, D = 37, 8192, 128                                   # 37 heads x 64 k-tiles = 16 waves of 148 CTAs
    q, k, v, dO = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(4)]
    qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
    doi, sdo = ops.quant_block(dO, 128)
    O, lse16, lse32 = ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D)
    delta = ops.bwd_delta(dO.view(-1, D), O)
    for _ in range(3):
        ops.int8_bwd_prequant(qi, ki, vi, doi, sq, sk, sv, sdo, lse32, delta, None, BH, S, D)
elif which == "fp4_fwd":
    from quantizedattention_b200 import attention_fp4 as F4
    q, k, v = [torch.randn(1, 37, 8192, 128, device="cuda", dtype=torch.float16) for _ in range(3)]   # 16 waves of 148 CTAs
    o = F4.quantise_fp4(q, k, v)
    for _ in range(3):
        F4.fp4_fwd_prequant(o, variant=int(__import__('os').environ.get('QA_FP4_VARIANT', '0')))
elif which == "fp4_quant":
    from quantizedattention_b200 import attention_fp4 as F4
    q, k, v = [torch.randn(2, 32, 8192, 128, device="cuda", dtype=torch.float16) for _ in range(3)]
    for _ in range(3):
        F4.quantise_fp4(q, k, v)
torch.cuda.synchronize()
print("ok")
