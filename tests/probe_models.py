"""Host-side models of the shared-memory / TMEM layouts the tcgen05 kernels assume.  Used by
tests/test_probe_gpu.py to check the models against the B200 (one MMA tile at a time)."""
import ctypes

import numpy as np
import torch

from quantizedattention_b200 import _lib

SW_BITS = {0: 0, 6: 1, 4: 2, 2: 3}       # UMMA layout type -> swizzle bits (none, 32B, 64B, 128B)
TMA_SWZ = {0: 0, 6: 1, 4: 2, 2: 3}       # UMMA layout type -> qa_make_tmap swizzle code


def swizzle_offsets(off, bits):
    """Swizzle<bits,4,3>: XOR address bits [4,4+bits) with bits [7,7+bits)."""
    return off ^ (((off >> 7) & ((1 << bits) - 1)) << 4)


def image_rows(mat_bytes: np.ndarray, layout: int) -> np.ndarray:
    """mat_bytes: [rows, row_bytes] uint8 with row_bytes == swizzle span (or any for no swizzle).
    Returns the linear shared-memory image: row r at r*row_bytes, 16-byte chunks XOR-swizzled."""
    rows, rb = mat_bytes.shape
    lin = np.arange(rows * rb, dtype=np.int64)
    dst = swizzle_offsets(lin, SW_BITS[layout])
    img = np.zeros(rows * rb, dtype=np.uint8)
    img[dst] = mat_bytes.reshape(-1)
    return img


def idesc(c_fmt, a_fmt, b_fmt, a_major, b_major, M, N):
    return (c_fmt << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run_mma(a_img, b_img, n_cols, *, a_lbo=16, a_sbo=1024, a_layout=2, a_kstep=32, b_lbo=16, b_sbo=1024, b_layout=2,
            b_kstep=32, idesc_v=0, kind=1, n_mma=4, a_in_tmem=0, a_tmem_cols=0, a_tmem_kstep_cols=8):
    L = _lib.dev_lib()
    a = torch.from_numpy(np.ascontiguousarray(a_img).view(np.uint8).reshape(-1)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(b_img).view(np.uint8).reshape(-1)).cuda()
    pad = lambda t: torch.cat([t, torch.zeros((-t.numel()) % 16, dtype=torch.uint8, device="cuda")])
    a, b = pad(a), pad(b)
    d = torch.zeros((128, n_cols), dtype=torch.int32, device="cuda")
    rc = L.qa_probe_mma(_lib.ptr(a), a.numel(), _lib.ptr(b), b.numel(), _lib.ptr(d), a_lbo, a_sbo, a_layout, a_kstep,
                        b_lbo, b_sbo, b_layout, b_kstep, ctypes.c_uint(idesc_v), kind, n_mma, n_cols, a_in_tmem,
                        a_tmem_cols, a_tmem_kstep_cols, _lib.cur_stream())
    _lib.check(rc, "qa_probe_mma", L)
    torch.cuda.synchronize()
    return d.cpu()


def run_tma(src: torch.Tensor, elem_bytes, dims, strides_bytes, box, swizzle, coords):
    L = _lib.dev_lib()
    rank = len(dims)
    nbytes = elem_bytes * int(np.prod(box))
    out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    U64 = ctypes.c_ulonglong * 3
    U32 = ctypes.c_uint * 3
    I32 = ctypes.c_int * 3
    d = U64(*(list(dims) + [1] * (3 - rank)))
    s = U64(*(list(strides_bytes) + [0] * (3 - len(strides_bytes))))
    b = U32(*(list(box) + [1] * (3 - rank)))
    c = I32(*(list(coords) + [0] * (3 - rank)))
    rc = L.qa_probe_tma(_lib.ptr(src), elem_bytes, rank, d, s, b, swizzle, c, _lib.ptr(out), _lib.cur_stream())
    _lib.check(rc, "qa_probe_tma", L)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ------------------------------------------------------------------ block-scaled (NVFP4) probe
E2M1 = np.array([0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6], dtype=np.float32)


def idesc_bs(a_fmt, b_fmt, M, N, sf_fmt, a_major=0, b_major=0, a_sf=0, b_sf=0):
    """Block-scaled instruction descriptor (csrc/qa_ptx.cuh umma_idesc_bs)."""
    return ((b_sf << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | (sf_fmt << 23)
            | ((M >> 4) << 24) | (a_sf << 29))


def pack_nibbles(codes):
    """[rows, K] e2m1 codes -> [rows, K/2] bytes, element 2i in the low nibble."""
    return (codes[:, 0::2] | (codes[:, 1::2] << 4)).astype(np.uint8)


def sf_atoms(sf):
    """[128, 4 * n] ue4m3 bytes -> n atoms of 512 B: byte 16 * (r % 32) + 4 * (r / 32) + s = scale of row r, block 4k + s."""
    n = sf.shape[1] // 4
    return np.ascontiguousarray(sf.reshape(4, 32, n, 4).transpose(2, 1, 0, 3)).reshape(-1)


def run_mma_bs(a_img, b_img, sfa_img, sfb_img, n_cols, n_mma, idesc_v, *, kind=0, a_in_tmem=0, a_tmem_cols=0, a_kcols=8, lay=4,
               sbo=512, kstep=32, sfb_col_offset=0):
    L = _lib.dev_lib()
    g = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).cuda()
    a, b, fa, fb = g(a_img), g(b_img), g(sfa_img), g(sfb_img)
    d = torch.zeros((128, n_cols), dtype=torch.float32, device="cuda")
    rc = L.qa_probe_mma_bs(_lib.ptr(a), a.numel(), _lib.ptr(b), b.numel(), _lib.ptr(fa), fa.numel(), _lib.ptr(fb), fb.numel(),
                           _lib.ptr(d), 16, sbo, lay, kstep, 16, sbo, lay, kstep, ctypes.c_uint(idesc_v), kind, n_mma, n_cols, 4, 4 | (sfb_col_offset << 8),
                           a_in_tmem, a_tmem_cols, a_kcols, _lib.cur_stream())
    _lib.check(rc, "qa_probe_mma_bs", L)
    torch.cuda.synchronize()
    return d.cpu().numpy()
