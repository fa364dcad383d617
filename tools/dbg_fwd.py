"""Development: compare (m, l, O_acc) of int8 forward kernel variants against an eager emulation, row by row."""
import math
import sys

import torch

sys.path.insert(0, ".")
from oracle import int8_ref as R  # noqa: E402
from quantizedattention_b200 import ops  # noqa: E402


def main(scale=3.0, seed=4, shape=(1, 2, 1024, 128)):
    g = torch.Generator().manual_seed(seed)
    q, k, v = [(torch.randn(shape, generator=g) * scale).to(torch.float16) for _ in range(3)]
    B, H, S, D = shape
    N = B * H * S
    q_i8, sq = R.quant_block(q.reshape(N, D), 128)
    k_i8, sk = R.quant_block(k.reshape(N, D), 128)
    G, L = B * H, S
    qg, kg = q_i8.view(G, L, D), k_i8.view(G, L, D)
    sq_rows = sq.view(G, L // 128).repeat_interleave(128, dim=1)[..., None].float()
    sk_g = sk.view(G, L // 128)
    qk_scale = (1 / math.sqrt(D)) * R.LOG2E
    l = torch.full((G, L, 1), 1.0)
    m = torch.full((G, L, 1), float("-inf"), dtype=torch.float16)
    for j in range(L // 128):
        acc = R._imm(qg, kg[:, j * 128:(j + 1) * 128].transpose(1, 2))
        S16 = (acc.float() * sq_rows * sk_g[:, j].view(G, 1, 1).float() * qk_scale).half()
        mn = torch.max(m, torch.amax(S16, -1, keepdim=True))
        l = l * torch.exp2((m - mn).float()) + torch.exp2((S16 - mn).float()).sum(-1, keepdim=True)
        m = mn
    qi, sqd = ops.quant_block(q.cuda(), 128); ki, skd = ops.quant_block(k.cuda(), 128); vi, svd = ops.quant_block(v.cuda(), 128)
    for ns in (2, 0):
        oacc, mm, ll = ops.int8_fwd_prequant(qi, ki, vi, sqd, skd, svd, G, S, S, D, 128, 128, nsplit=ns, ring_state=True)
        torch.cuda.synchronize()
        mm, ll = mm.cpu(), ll.cpu()
        dm = (mm - m.flatten().float()).abs()
        dl = ((ll - l.flatten()) / l.flatten()).abs()
        i = int(dl.argmax())
        print(f"nsplit{ns}: max|dm| {dm.max().item():.3e}  max rel dl {dl.max().item():.3e} at row {i}: l_k {ll[i].item():.6f} l_ref "
              f"{l.flatten()[i].item():.6f} m_k {mm[i].item()} m_ref {m.flatten()[i].item()}  rows with rel dl > 1e-3: {(dl > 1e-3).sum().item()}")
        bad = torch.nonzero(dl > 1e-3).flatten()[:20].tolist()
        print("   bad rows:", bad, [round(float(dl[b]), 4) for b in bad])


if __name__ == "__main__":
    main()
    main(1.0, 1, (1, 2, 256, 128))
