"""Development: per-CTA globaltimer stamps of the warp-specialised bf16 backward (libqattn_dev.so).  Prints the mean
duration of each phase of a CTA and the gap between consecutive CTAs on one SM.
Slots: 0 entry, 1 setup done, 2 K/V/Q landed, 3 P(0) ready, 4 dS(0) ready, 5 P(1) ready, 6 P(2) ready, 7 S(0) seen by compute,
8 P(0) phase end, 9 dS(0) phase end, 10 dS(last) end, 11 accumulators complete, 12 epilogue stores issued, 13 dQ reds issued,
14 exit, 15 SM id."""
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
    dO = torch.randn(B, H, S, D, device="cuda")
    O, lse = ops.bf16_fwd(q, k, v, bool(causal))
    for _ in range(2):
        ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO)
    n_cta = B * H * (S // 128)
    buf = torch.zeros(n_cta, 64, dtype=torch.int64, device="cuda")
    L = _lib.dev_lib()
    L.qa_debug_set_bf16_bwd_timeline(_lib.ptr(buf))
    ops.bf16_bwd(q, k, v, O, lse, bool(causal), dO)
    torch.cuda.synchronize()
    L.qa_debug_set_bf16_bwd_timeline(None)
    t = buf.cpu().view(B * H, S // 128, 64)
    t0 = t[..., 0].min()
    print(f"B={B} H={H} S={S} causal={causal}: kernel span {(t[..., 14].max() - t0).item() / 1e3:.1f} us")
    names = ["setup (entry -> barriers/TMEM ready)", "K/V/Q(0) landed", "P(0) ready at the MMA warp", "dS(0) ready", "P(1) ready", "P(2) ready"]
    for jsel in ([0, 8, 16, 24, 31] if causal else [0]):
        c = t[:, jsel].double()
        tiles = (S // 128 - jsel) if causal else S // 128
        d = lambda a, b: ((c[:, a] - c[:, b]).mean().item() / 1e3)
        print(f"  key tile {jsel} ({tiles} query tiles): total {d(14, 0):.2f} us | setup {d(1, 0):.2f} | loads {d(2, 1):.2f} | "
              f"S(0) seen {d(7, 2):.2f} | P(0) phase {d(8, 7):.2f} | P(0)->MMA {d(3, 8):.2f} | dS(0) ready {d(4, 3):.2f} | "
              f"P(1) {d(5, 4):.2f} | P(2) {d(6, 5):.2f} | steady per tile {(d(10, 6) / max(tiles - 3, 1)):.2f} | "
              f"last dS -> acc complete {d(11, 10):.2f} | epilogue {d(12, 11):.2f} | exit after epilogue {d(14, 12):.2f} | reds done before exit {d(14, 13):.2f}")
    c = t[:, 0].double()
    e = lambda a: (c[:, a] - c[:, 16]).mean().item() / 1e3
    print("  steady state, query tile 4 of key tile 0, us after P_READY(4) reached the MMA warp:")
    print(f"    MMA warp : dV(4)+S(5) issued {e(17):.2f} | dS(4) ready seen {e(18):.2f} | dQ(4)+dK(4) issued {e(19):.2f} | dQ(4) drained seen {e(20):.2f} | P_READY(5) seen {e(21):.2f}")
    print(f"    compute  : S(4) seen {e(22):.2f} | P(4) written {e(23):.2f} | dP(4) seen {e(24):.2f} | dS(4) written {e(25):.2f}")
    print(f"    dQ drain : dQ(4) seen {e(26):.2f} | TMEM read done {e(27):.2f} | reds issued {e(28):.2f}")
    for cw in range(8):
        b = 32 + cw * 4
        print(f"    compute warp {cw + 4} (SMSP {cw % 4}): S seen {e(b):.2f} | P written {e(b + 1):.2f} (phase {e(b + 1) - e(b):.2f}) | dP seen {e(b + 2):.2f} | "
              f"dS written {e(b + 3):.2f} (phase {e(b + 3) - e(b + 2):.2f})")
    # gaps between consecutive CTAs of one SM
    flat = t.view(-1, 64)
    gaps = []
    for sm in flat[:, 15].unique().tolist():
        m = flat[flat[:, 15] == sm]
        m = m[m[:, 0].argsort()]
        if len(m) > 1:
            gaps.append((m[1:, 0] - m[:-1, 14]).double())
    g = torch.cat(gaps)
    print(f"  exit -> next CTA entry on the same SM: mean {g.mean().item() / 1e3:.2f} us, median {g.median().item() / 1e3:.2f}, max {g.max().item() / 1e3:.2f}")
    busy = (flat[:, 14] - flat[:, 0]).double().sum().item() / 148 / 1e3
    print(f"  mean busy time per SM {busy:.1f} us")


if __name__ == "__main__":
    main()
    main(1, 37, 2048, 0)
