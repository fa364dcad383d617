"""Development: pins the NVFP4 block-scaled tcgen05.mma conventions on the B200 (operand nibble order, 64-byte-swizzled K-major
rows, scale-factor atom layout, TS-mode A) through qa_probe_mma_bs; prints the error of each hypothesis."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from quantizedattention_b200 import _lib  # noqa: E402
from probe_models import image_rows  # noqa: E402

E2M1 = np.array([0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6], dtype=np.float32)


def idesc_bs(a_fmt, b_fmt, M, N, sf_fmt, a_major=0, b_major=0, a_sf=0, b_sf=0):
    return (b_sf << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | (sf_fmt << 23) | ((M >> 4) << 24) | (a_sf << 29)


def pack_nibbles(codes, low_first=True):
    lo, hi = (codes[:, 0::2], codes[:, 1::2]) if low_first else (codes[:, 1::2], codes[:, 0::2])
    return (lo | (hi << 4)).astype(np.uint8)


def sf_atoms(sf):                       # sf: [128, 4 * n_mma] uint8 -> n_mma atoms of 512 B
    rows, n = sf.shape
    out = np.zeros((n // 4, 512), dtype=np.uint8)
    for k in range(n // 4):
        for r in range(rows):
            for s in range(4):
                out[k, 16 * (r % 32) + 4 * (r // 32) + s] = sf[r, 4 * k + s]
    return out.reshape(-1)


def e4m3_to_f32(b):
    return torch.from_numpy(b.copy()).view(torch.float8_e4m3fn).float().numpy()


def run(a_img, b_img, sfa_img, sfb_img, n_cols, n_mma, idesc, kind=0, a_in_tmem=0, a_tmem_cols=0, a_kcols=8, lay=4, sbo=512, kstep=32):
    L = _lib.dev_lib()
    g = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).cuda()
    a, b, fa, fb = g(a_img), g(b_img), g(sfa_img), g(sfb_img)
    d = torch.zeros((128, n_cols), dtype=torch.float32, device="cuda")
    rc = L.qa_probe_mma_bs(_lib.ptr(a), a.numel(), _lib.ptr(b), b.numel(), _lib.ptr(fa), fa.numel(), _lib.ptr(fb), fb.numel(), _lib.ptr(d),
                           16, sbo, lay, kstep, 16, sbo, lay, kstep, ctypes.c_uint(idesc), kind, n_mma, n_cols, 4, 4, a_in_tmem, a_tmem_cols,
                           a_kcols, _lib.cur_stream())
    _lib.check(rc, "qa_probe_mma_bs", L)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def main():
    rng = np.random.default_rng(0)
    n_mma, N = 2, 128
    K = 64 * n_mma
    ac, bc = rng.integers(0, 16, (128, K)), rng.integers(0, 16, (N, K))
    sfa = rng.integers(0x28, 0x48, (128, K // 16)).astype(np.uint8)           # ue4m3 codes around 1.0
    sfb = rng.integers(0x28, 0x48, (N, K // 16)).astype(np.uint8)
    A, B = E2M1[ac], E2M1[bc]
    fa, fb = e4m3_to_f32(sfa), e4m3_to_f32(sfb)
    ref = np.zeros((128, N), dtype=np.float64)
    for blk in range(K // 16):
        sl = slice(16 * blk, 16 * blk + 16)
        ref += (A[:, sl].astype(np.float64) @ B[:, sl].astype(np.float64).T) * fa[:, blk:blk + 1] * fb[:, blk][None, :]
    ref_noscale = A.astype(np.float64) @ B.astype(np.float64).T
    ide = idesc_bs(1, 1, 128, N, 0)
    for low_first in (True, False):
        a_img = image_rows(pack_nibbles(ac, low_first), 4)
        b_img = image_rows(pack_nibbles(bc, low_first), 4)
        got = run(a_img, b_img, sf_atoms(sfa), sf_atoms(sfb), N, n_mma, ide)
        e = np.abs(got - ref).max() / np.abs(ref).max()
        e0 = np.abs(got - ref_noscale).max() / np.abs(ref_noscale).max()
        print(f"SS low_nibble_first={low_first}: rel err vs scaled ref {e:.3e}, vs unscaled {e0:.3e}; sample {got[0, :4]} ref {ref[0, :4]}")
    # TS mode: A (packed e2m1) in TMEM, 8 elements per 32-bit column, K = 64 -> 8 columns per MMA
    a_words = pack_nibbles(ac, True).view(np.uint32).reshape(128, K // 8)
    got = run(a_words, image_rows(pack_nibbles(bc, True), 4), sf_atoms(sfa), sf_atoms(sfb), N, n_mma, ide, a_in_tmem=1, a_tmem_cols=K // 8, a_kcols=8)
    print(f"TS: rel err {np.abs(got - ref).max() / np.abs(ref).max():.3e}; sample {got[0, :4]}")


if __name__ == "__main__":
    main()
