import sys, torch
sys.path.insert(0, ".")
from oracle import bf16_ref
from oracle.baseline import baseline_pytorch_attention
from quantizedattention_b200 import ops
g = torch.Generator().manual_seed(1)
for shape in [(1, 4, 1024, 128), (1, 2, 2048, 64)]:
    q, k, v = [torch.randn(shape, generator=g) for _ in range(3)]
    q, k, v = q.half(), k.half(), v.bfloat16()
    O, lse = ops.bf16_fwd(q.cuda(), k.cuda(), v.cuda(), False)
    Or, lser = bf16_ref.bf16_fwd(q, k, v, False, tile_k=128, mode="contract")
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), shape[3], False)
    d1 = (O.cpu() - Or); d2 = (O.cpu() - base); d3 = (Or - base)
    print(shape, "vs oracle max %.2e mse %.2e | vs fp32 max %.2e mse %.2e | oracle vs fp32 max %.2e mse %.2e | lse %.2e" % (
        d1.abs().max(), (d1 ** 2).mean(), d2.abs().max(), (d2 ** 2).mean(), d3.abs().max(), (d3 ** 2).mean(), (lse.cpu() - lser).abs().max()))
