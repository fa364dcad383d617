"""Probe cases (one MMA tile each) for tests/test_probe_gpu.py.  Run as a script with case names: each case
prints one JSON line; the parent test restarts the script after a case that faults the CUDA context."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import probe_models as pm  # noqa: E402

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _i8(rng, r, c):
    return rng.integers(-127, 128, size=(r, c), dtype=np.int8)


def _u8(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], -1)


def _bf(rng, r, c, absval=False):
    x = rng.standard_normal((r, c)).astype(np.float32)
    if absval:
        x = np.abs(x)
    return torch.from_numpy(x).to(torch.bfloat16)


def _bimg(t, lay):
    return pm.image_rows(t.contiguous().view(torch.int16).numpy().view(np.uint8).reshape(t.shape[0], -1), lay)


def _eq(got, exp):
    return bool((got.numpy() == exp).all())


def _err(got, exp):
    return float(np.abs(got.numpy().view(np.float32) - exp).max())


# ------------------------------------------------------------------ int8
@case
def i8_kmajor_sw128():
    rng = np.random.default_rng(1)
    A, B = _i8(rng, 128, 128), _i8(rng, 128, 128)
    got = pm.run_mma(pm.image_rows(_u8(A), 2), pm.image_rows(_u8(B), 2), 128, idesc_v=pm.idesc(2, 1, 1, 0, 0, 128, 128),
                     kind=1, n_mma=4)
    return _eq(got, A.astype(np.int32) @ B.astype(np.int32).T)


@case
def i8_kmajor_sw64():
    rng = np.random.default_rng(2)
    A, B = _i8(rng, 128, 64), _i8(rng, 128, 64)
    got = pm.run_mma(pm.image_rows(_u8(A), 4), pm.image_rows(_u8(B), 4), 128, idesc_v=pm.idesc(2, 1, 1, 0, 0, 128, 128),
                     kind=1, n_mma=2, a_sbo=512, a_layout=4, b_sbo=512, b_layout=4)
    return _eq(got, A.astype(np.int32) @ B.astype(np.int32).T)


def _pv(lbo, sbo):
    rng = np.random.default_rng(3)
    P, V = rng.integers(0, 128, size=(128, 128), dtype=np.int8), _i8(rng, 128, 128)
    got = pm.run_mma(pm.image_rows(_u8(P), 2), pm.image_rows(_u8(V), 2), 128, idesc_v=pm.idesc(2, 1, 1, 0, 1, 128, 128),
                     kind=1, n_mma=4, b_lbo=lbo, b_sbo=sbo, b_kstep=4096)
    return _eq(got, P.astype(np.int32) @ V.astype(np.int32))


@case
def i8_Bmn_sw128_lbo16_sbo1024():
    return _pv(16, 1024)


@case
def i8_Bmn_sw128_lbo1024_sbo16():
    return _pv(1024, 16)


@case
def i8_Bmn_sw128_lbo1024_sbo1024():
    return _pv(1024, 1024)


def _pv64(lbo, sbo):
    rng = np.random.default_rng(4)
    P, V = rng.integers(0, 128, size=(128, 128), dtype=np.int8), _i8(rng, 128, 64)
    got = pm.run_mma(pm.image_rows(_u8(P), 2), pm.image_rows(_u8(V), 4), 64, idesc_v=pm.idesc(2, 1, 1, 0, 1, 128, 64),
                     kind=1, n_mma=4, b_lbo=lbo, b_sbo=sbo, b_layout=4, b_kstep=2048)
    return _eq(got, P.astype(np.int32) @ V.astype(np.int32))


@case
def i8_Bmn_sw64_lbo16_sbo512():
    return _pv64(16, 512)


@case
def i8_Bmn_sw64_lbo512_sbo16():
    return _pv64(512, 16)


@case
def i8_Bmn_sw64_lbo512_sbo512():
    return _pv64(512, 512)


def _at(lbo, sbo):
    rng = np.random.default_rng(5)
    Pm, dO = _i8(rng, 128, 128), _i8(rng, 128, 128)
    got = pm.run_mma(pm.image_rows(_u8(Pm), 2), pm.image_rows(_u8(dO), 2), 128, idesc_v=pm.idesc(2, 1, 1, 1, 1, 128, 128),
                     kind=1, n_mma=4, a_lbo=lbo, a_sbo=sbo, a_kstep=4096, b_lbo=lbo, b_sbo=sbo, b_kstep=4096)
    return _eq(got, Pm.astype(np.int32).T @ dO.astype(np.int32))


@case
def i8_Amn_sw128_lbo16_sbo1024():
    return _at(16, 1024)


@case
def i8_Amn_sw128_lbo1024_sbo1024():
    return _at(1024, 1024)


@case
def i8_TS_packed4():
    rng = np.random.default_rng(6)
    P, V = rng.integers(0, 128, size=(128, 128), dtype=np.int8), _i8(rng, 128, 128)
    a_t = _u8(P).reshape(128, 32, 4).copy().view(np.uint32).reshape(128, 32)
    got = pm.run_mma(a_t, pm.image_rows(_u8(V), 2), 128, idesc_v=pm.idesc(2, 1, 1, 0, 1, 128, 128), kind=1, n_mma=4,
                     b_lbo=16, b_sbo=1024, b_kstep=4096, a_in_tmem=1, a_tmem_cols=32, a_tmem_kstep_cols=8)
    return _eq(got, P.astype(np.int32) @ V.astype(np.int32))


# ------------------------------------------------------------------ 16-bit
@case
def f16_kmajor_sw128():
    rng = np.random.default_rng(7)
    f16 = lambda r, c: (rng.standard_normal((r, c)) * 0.5).astype(np.float16)
    A, B = f16(128, 64), f16(128, 64)
    got = pm.run_mma(pm.image_rows(_u8(A), 2), pm.image_rows(_u8(B), 2), 128, idesc_v=pm.idesc(1, 0, 0, 0, 0, 128, 128),
                     kind=0, n_mma=4)
    return _err(got, A.astype(np.float32) @ B.astype(np.float32).T)


def _pv16(lbo, sbo):
    rng = np.random.default_rng(8)
    P, V = _bf(rng, 128, 64, True), _bf(rng, 64, 64)
    got = pm.run_mma(_bimg(P, 2), _bimg(V, 2), 64, idesc_v=pm.idesc(1, 1, 1, 0, 1, 128, 64), kind=0, n_mma=4, b_lbo=lbo,
                     b_sbo=sbo, b_kstep=2048)
    return _err(got, (P.float() @ V.float()).numpy())


@case
def bf16_Bmn_sw128_lbo16_sbo1024():
    return _pv16(16, 1024)


@case
def bf16_Bmn_sw128_lbo1024_sbo1024():
    return _pv16(1024, 1024)


def _pv16_2(lbo, sbo):
    rng = np.random.default_rng(9)
    P, V2 = _bf(rng, 128, 64, True), _bf(rng, 64, 128)
    v_img = np.concatenate([_bimg(V2[:, :64], 2), _bimg(V2[:, 64:], 2)])
    got = pm.run_mma(_bimg(P, 2), v_img, 128, idesc_v=pm.idesc(1, 1, 1, 0, 1, 128, 128), kind=0, n_mma=4, b_lbo=lbo,
                     b_sbo=sbo, b_kstep=2048)
    return _err(got, (P.float() @ V2.float()).numpy())


@case
def bf16_Bmn_2atoms_lbo8192_sbo1024():
    return _pv16_2(8192, 1024)


@case
def bf16_Bmn_2atoms_lbo1024_sbo8192():
    return _pv16_2(1024, 8192)


@case
def bf16_TS_packed2():
    rng = np.random.default_rng(8)
    P, V = _bf(rng, 128, 64, True), _bf(rng, 64, 64)
    a_t = P.view(torch.int16).numpy().view(np.uint16).reshape(128, 32, 2).copy().view(np.uint32).reshape(128, 32)
    got = pm.run_mma(a_t, _bimg(V, 2), 64, idesc_v=pm.idesc(1, 1, 1, 0, 1, 128, 64), kind=0, n_mma=4, b_lbo=16,
                     b_sbo=1024, b_kstep=2048, a_in_tmem=1, a_tmem_cols=32, a_tmem_kstep_cols=8)
    return _err(got, (P.float() @ V.float()).numpy())


def _at16(lbo, sbo):
    rng = np.random.default_rng(10)
    Pq, dO = _bf(rng, 64, 128), _bf(rng, 64, 64)          # [q][keys], [q][D]
    a_img = np.concatenate([_bimg(Pq[:, :64], 2), _bimg(Pq[:, 64:], 2)])
    got = pm.run_mma(a_img, _bimg(dO, 2), 64, idesc_v=pm.idesc(1, 1, 1, 1, 1, 128, 64), kind=0, n_mma=4, a_lbo=lbo,
                     a_sbo=sbo, a_kstep=2048, b_lbo=16, b_sbo=1024, b_kstep=2048)
    return _err(got, (Pq.float().T @ dO.float()).numpy())


@case
def bf16_Amn_2atoms_lbo8192_sbo1024():
    return _at16(8192, 1024)


@case
def bf16_Amn_2atoms_lbo1024_sbo8192():
    return _at16(1024, 8192)


# ------------------------------------------------------------------ NVFP4 block-scaled (kind::mxf4nvf4.block_scale.block16)
def _nvf4(ts):
    rng = np.random.default_rng(11)
    n_mma, N = 2, 128
    K = 64 * n_mma
    ac, bc = rng.integers(0, 16, (128, K)), rng.integers(0, 16, (N, K))
    sfa = rng.integers(0x28, 0x48, (128, K // 16)).astype(np.uint8)           # ue4m3 codes around 1.0
    sfb = rng.integers(0x28, 0x48, (N, K // 16)).astype(np.uint8)
    A, B = pm.E2M1[ac].astype(np.float64), pm.E2M1[bc].astype(np.float64)
    f = lambda b: torch.from_numpy(b.copy()).view(torch.float8_e4m3fn).float().numpy().astype(np.float64)
    fa, fb = f(sfa), f(sfb)
    ref = np.zeros((128, N))
    for blk in range(K // 16):
        sl = slice(16 * blk, 16 * blk + 16)
        ref += (A[:, sl] @ B[:, sl].T) * fa[:, blk:blk + 1] * fb[:, blk][None, :]
    ide = pm.idesc_bs(1, 1, 128, N, 0)
    b_img = pm.image_rows(pm.pack_nibbles(bc), 4)                            # rows of 64 bytes, 64-byte swizzle
    if ts:                                                                   # A from TMEM: 8 e2m1 per 32-bit column
        a_words = pm.pack_nibbles(ac).view(np.uint32).reshape(128, K // 8)
        got = pm.run_mma_bs(a_words, b_img, pm.sf_atoms(sfa), pm.sf_atoms(sfb), N, n_mma, ide, a_in_tmem=1, a_tmem_cols=K // 8)
    else:
        got = pm.run_mma_bs(pm.image_rows(pm.pack_nibbles(ac), 4), b_img, pm.sf_atoms(sfa), pm.sf_atoms(sfb), N, n_mma, ide)
    return float(np.abs(got - ref).max() / np.abs(ref).max())


@case
def nvf4_blockscaled_n64():
    """N = 64 (a 64-key step): the B scale factors are the first two columns (rows 0..63) of their own atom."""
    rng = np.random.default_rng(12)
    n_mma, N = 2, 64
    K = 64 * n_mma
    ac, bc = rng.integers(0, 16, (128, K)), rng.integers(0, 16, (N, K))
    sfa = rng.integers(0x28, 0x48, (128, K // 16)).astype(np.uint8)
    sfb = rng.integers(0x28, 0x48, (N, K // 16)).astype(np.uint8)
    A, B = pm.E2M1[ac].astype(np.float64), pm.E2M1[bc].astype(np.float64)
    f = lambda b: torch.from_numpy(b.copy()).view(torch.float8_e4m3fn).float().numpy().astype(np.float64)
    fa, fb = f(sfa), f(sfb)
    ref = np.zeros((128, N))
    for blk in range(K // 16):
        sl = slice(16 * blk, 16 * blk + 16)
        ref += (A[:, sl] @ B[:, sl].T) * fa[:, blk:blk + 1] * fb[:, blk][None, :]
    sfb_pad = np.concatenate([sfb, np.zeros((64, K // 16), dtype=np.uint8)])
    got = pm.run_mma_bs(pm.image_rows(pm.pack_nibbles(ac), 4), pm.image_rows(pm.pack_nibbles(bc), 4), pm.sf_atoms(sfa), pm.sf_atoms(sfb_pad), N,
                        n_mma, pm.idesc_bs(1, 1, 128, N, 0))
    return float(np.abs(got - ref).max() / np.abs(ref).max())


@case
def nvf4_blockscaled_n64_upper_half():
    """N = 64 on keys 64..127 of a 128-key tile: B's scales are columns 2, 3 of the tile's 128-row atom (TMEM address + 2)."""
    rng = np.random.default_rng(13)
    n_mma, N = 2, 64
    K = 64 * n_mma
    ac, bc = rng.integers(0, 16, (128, K)), rng.integers(0, 16, (128, K))
    sfa = rng.integers(0x28, 0x48, (128, K // 16)).astype(np.uint8)
    sfb = rng.integers(0x28, 0x48, (128, K // 16)).astype(np.uint8)
    A, B = pm.E2M1[ac].astype(np.float64), pm.E2M1[bc[64:]].astype(np.float64)
    f = lambda b: torch.from_numpy(b.copy()).view(torch.float8_e4m3fn).float().numpy().astype(np.float64)
    fa, fb = f(sfa), f(sfb[64:])
    ref = np.zeros((128, N))
    for blk in range(K // 16):
        sl = slice(16 * blk, 16 * blk + 16)
        ref += (A[:, sl] @ B[:, sl].T) * fa[:, blk:blk + 1] * fb[:, blk][None, :]
    got = pm.run_mma_bs(pm.image_rows(pm.pack_nibbles(ac), 4), pm.image_rows(pm.pack_nibbles(bc[64:]), 4), pm.sf_atoms(sfa), pm.sf_atoms(sfb), N,
                        n_mma, pm.idesc_bs(1, 1, 128, N, 0), sfb_col_offset=2)
    return float(np.abs(got - ref).max() / np.abs(ref).max())


@case
def nvf4_blockscaled_ss_sw64():
    return _nvf4(False)


@case
def nvf4_blockscaled_ts():
    return _nvf4(True)


if __name__ == "__main__":
    for name in sys.argv[1:]:
        r = CASES[name]()
        print("PROBE " + json.dumps({"case": name, "result": r}), flush=True)
