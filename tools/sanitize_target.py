"""Small shapes of every fused kernel for compute-sanitizer (one tool per run, e.g.
`compute-sanitizer --tool racecheck python tools/sanitize_target.py int8`).  Prints "ok" when the kernels ran."""
import sys

import torch

sys.path.insert(0, ".")
from quantizedattention_b200 import attention_bf16 as Bf  # noqa: E402
from quantizedattention_b200 import attention_int8 as A  # noqa: E402
from quantizedattention_b200 import attention_jvp as J  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "int8"
g = torch.Generator().manual_seed(0)
if which == "int8":
    q, k, v, dO = [torch.randn(1, 2, 256, 128, generator=g).half().cuda() for _ in range(4)]
    qr, kr, vr = [t.requires_grad_() for t in (q, k, v)]
    A.sage_attention_3_int8(qr, kr, vr).backward(dO)
elif which == "bf16":
    q, k, v, dO = [torch.randn(1, 2, 256, 128, generator=g).cuda() for _ in range(4)]
    qr, kr, vr = q.half().requires_grad_(), k.half().requires_grad_(), v.bfloat16().requires_grad_()
    Bf.flash_atten_2_bf16(qr, kr, vr, True).backward(dO)
else:
    t6 = [torch.randn(1, 2, 256, 64, generator=g).cuda() for _ in range(6)]
    J.helion_attention_jvp_forward_fp32(*t6)
torch.cuda.synchronize()
print("ok")
