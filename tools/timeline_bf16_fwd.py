"""Development: per-CTA globaltimer stamps of the two-tile bf16 forward (libqattn_dev.so).
Slots: 0 entry, 1 setup done, 2/3 first S seen by tile A/B, 4/5 step 4 seen, 6/7 softmax loop end, 8/9 last PV complete,
10/11 epilogue end, 12 exit, 15 SM id."""
import os
import sys

os.environ["QA_DEV_LIB"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from quantizedattention_b200 import _lib, ops  # noqa: E402


def main(B=4, H=16, S=4096, causal=1):
    D = 128
    torch.manual_seed(0)
    q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.bf16_fwd(q, k, v, bool(causal))
    npair = S // 256
    buf = torch.zeros(B * H * npair, 16, dtype=torch.int64, device="cuda")
    L = _lib.dev_lib()
    L.qa_debug_set_bf16_fwd_timeline(_lib.ptr(buf))
    ops.bf16_fwd(q, k, v, bool(causal))
    torch.cuda.synchronize()
    L.qa_debug_set_bf16_fwd_timeline(None)
    G = 16 if causal else 1                                # launch order: qa_group_order (csrc/qa_ptx.cuh)
    w = torch.arange(B * H * npair)
    g = w // (G * npair)
    rem = w - g * (G * npair)
    rank = rem // G
    t = buf.cpu().double()                                 # [blockIdx.x]; pair index pt = npair - 1 - rank
    t0 = t[..., 0].min()
    print(f"B={B} H={H} S={S} causal={causal}: kernel span {(t[..., 12].max() - t0).item() / 1e3:.1f} us")
    for bx in ([0, npair // 2, npair - 1] if causal else [0]):
        c = t[rank == bx]
        pt = npair - 1 - bx
        nk = (2 * pt + 2) if causal else S // 128
        d = lambda a, b: ((c[:, a] - c[:, b]).mean().item() / 1e3)
        print(f"  pair {pt} ({nk} k-tiles): total {d(12, 0):.2f} us | setup {d(1, 0):.2f} | first S (tile B) {d(3, 1):.2f} | step 4 {d(5, 3):.2f} | "
              f"steady per k-tile {(d(7, 5) / max(nk - 2, 1)):.2f} | loop end -> last PV {d(9, 7):.2f} | epilogue {d(11, 9):.2f} | exit {d(12, 11):.2f}")
    flat = t
    gaps = []
    for sm in flat[:, 15].unique().tolist():
        m = flat[flat[:, 15] == sm]
        m = m[m[:, 0].argsort()]
        if len(m) > 1:
            gaps.append(m[1:, 0] - m[:-1, 12])
    g = torch.cat(gaps)
    print(f"  exit -> next CTA entry on the same SM: mean {g.mean().item() / 1e3:.2f} us")
    fin = []
    for sm in flat[:, 15].unique().tolist():
        fin.append(flat[flat[:, 15] == sm][:, 12].max().item())
    print(f"  SM finish times: min {(min(fin) - t0.item()) / 1e3:.1f} us, max {(max(fin) - t0.item()) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
    main(4, 16, 4096, 0)
