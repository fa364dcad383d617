"""Development: per-item globaltimer stamps of the warp-specialised bf16 backward (libqattn_dev.so).
Buffer [item = head * key tiles + key tile][64]; slots: 0 item start at the MMA warp, 2 K/V/Q of the item landed, 3 P(0)
ready, 11 accumulators complete (seen by the drain warps), 12 dV / dK stores issued, 14 = 1 + the CTA that processed the item.
The per-tile stamps of the first (one item per CTA) version are kept in profiles/r02_bf16_bwd_timeline.txt."""
import os
import sys
from collections import defaultdict

os.environ["QA_DEV_LIB"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib, ops  # noqa: E402


def main(B=4, H=16, S=4096, causal=1):
    D = 128
    torch.manual_seed(0)
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B, H, S, D, device="cuda")
    O, lse = ops.bf16_fwd(q, k, v, bool(causal))
    for _ in range(2):
        ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO)
    n_items = B * H * (S // 128)
    buf = torch.zeros(n_items, 64, dtype=torch.int64, device="cuda")
    L = _lib.dev_lib()
    L.qa_debug_set_bf16_bwd_timeline(_lib.ptr(buf))
    ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO)
    torch.cuda.synchronize()
    L.qa_debug_set_bf16_bwd_timeline(None)
    t = buf.cpu()
    t0 = int(t[:, 0].min())
    per_cta = defaultdict(list)
    for i in range(n_items):
        per_cta[int(t[i, 14])].append(i)
    print(f"B={B} H={H} S={S} causal={causal}: {n_items} items processed by {len(per_cta)} CTAs; span {(int(t[:, 12].max()) - t0) / 1e3:.1f} us")
    gaps, firstp, loads, wb, n_per = [], [], [], [], []
    for cta, items in per_cta.items():
        items.sort(key=lambda i: int(t[i, 0]))
        n_per.append(len(items))
        for a, b in zip(items[:-1], items[1:]):
            gaps.append((int(t[b, 0]) - int(t[a, 11])) / 1e3)          # next item start relative to accumulators complete
            firstp.append((int(t[b, 3]) - int(t[a, 11])) / 1e3)        # first P of the next item
        for i in items:
            loads.append((int(t[i, 2]) - int(t[i, 0])) / 1e3)
            wb.append((int(t[i, 12]) - int(t[i, 11])) / 1e3)
    m = lambda x: sum(x) / max(len(x), 1)
    print(f"  items per CTA: min {min(n_per)}, max {max(n_per)}")
    print(f"  loads landed {m(loads):.2f} us after item start | dV / dK write-back {m(wb):.2f} us | next item starts {m(gaps):+.2f} us "
          f"after the accumulators of the previous one are complete, its first P is ready after {m(firstp):.2f} us")
    fin = [max(int(t[i, 12]) for i in items) for items in per_cta.values()]
    print(f"  CTA finish times: min {(min(fin) - t0) / 1e3:.1f} us, max {(max(fin) - t0) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
    main(1, 32, 8192, 0)
