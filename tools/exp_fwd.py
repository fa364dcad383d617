"""Development check of int8 forward kernel variants (nsplit selects the kernel): parity vs the oracle at small shapes,
then kernel timing at the cfg3 tile shape.  Usage: python tools/exp_fwd.py [BH] -> gpurun_out/exp_fwd.json"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from oracle import int8_ref  # noqa: E402
from quantizedattention_b200 import ops  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

VARIANTS = [int(x) for x in os.environ.get("QA_VARIANTS", "2,0").split(",")]


def parity(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    q, k, v = [(torch.randn(shape, generator=g) * scale).to(torch.float16) for _ in range(3)]
    B, H, S, D = shape
    ref = int8_ref.int8_fwd(q, k, v, 128, 128, per_head=True, return_lse32=True)
    qi, sq = ops.quant_block(q.cuda(), 128); ki, sk = ops.quant_block(k.cuda(), 128); vi, sv = ops.quant_block(v.cuda(), 128)
    out = {}
    for ns in VARIANTS:
        O, lse16, lse32 = ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, B * H, S, S, D, 128, 128, nsplit=ns)
        torch.cuda.synchronize()
        a, b = O.cpu().float().flatten(), ref[0].float().flatten()
        out[f"nsplit{ns}"] = {"max_abs": (a - b).abs().max().item(), "mse": ((a - b) ** 2).mean().item(),
                              "cos": torch.nn.functional.cosine_similarity(a, b, dim=0).item(),
                              "lse32_max": (lse32.cpu() - ref[10]).abs().max().item()}
    return out


def main():
    BH = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    res = {"parity": {}}
    for shape, seed, scale in [((1, 2, 256, 128), 1, 1.0), ((1, 8, 1024, 64), 2, 1.0), ((1, 2, 2048, 128), 3, 1.0),
                               ((1, 2, 1024, 128), 4, 3.0)]:
        res["parity"][f"{shape}x{scale}"] = parity(shape, seed, scale)
    for D in (128, 64):
        S = 8192
        torch.manual_seed(0)
        q, k, v = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(3)]
        qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
        for ns in VARIANTS:
            med, best = timeit(lambda: ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, 128, 128, nsplit=ns), warm=3, it=8)
            res[f"time_D{D}_nsplit{ns}"] = {"ms_med": med, "ms_best": best, "TOPS_med": 4 * BH * S * S * D / med / 1e9,
                                           "clk_per_tile_1965": med * 1e-3 * 1.965e9 / (BH * (S / 128) ** 2 / 148)}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/exp_fwd.json", "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
