"""Hardware probes: confirm on the B200 the TMA swizzle, UMMA shared-memory descriptor and TMEM operand
layouts assumed by the attention kernels (quantizedattention_b200/csrc/attn_*.cu).  Each case computes one
MMA tile and compares with exact integer / fp32 host math."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPORT = {}


def _save():
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_report.json", "w") as f:
        json.dump(REPORT, f, indent=1)


def _i8(rng, r, c):
    return rng.integers(-127, 128, size=(r, c), dtype=np.int8)


def test_tma_swizzle_models():
    import probe_models as pm
    rng = np.random.default_rng(0)
    src = torch.from_numpy(rng.integers(0, 256, size=(512, 256), dtype=np.uint8)).cuda()
    res = {}
    for layout, row_bytes in ((2, 128), (4, 64), (6, 32)):
        got = pm.run_tma(src, 1, [256, 512], [256], [row_bytes, 64], pm.TMA_SWZ[layout], [row_bytes, 128])
        tile = src.cpu().numpy()[128:192, row_bytes:2 * row_bytes]
        exp = pm.image_rows(tile, layout)
        res[f"sw{row_bytes}"] = bool((got == exp).all())
    # 3-D map [D, S, BH] as the attention kernels use it
    src3 = torch.from_numpy(rng.integers(0, 256, size=(3, 256, 128), dtype=np.uint8)).cuda()
    got = pm.run_tma(src3, 1, [128, 256, 3], [128, 256 * 128], [128, 128, 1], 3, [0, 128, 2])
    exp = pm.image_rows(src3.cpu().numpy()[2, 128:256, :], 2)
    res["sw128_3d"] = bool((got == exp).all())
    REPORT["tma"] = res
    _save()
    assert all(res.values()), res


def test_umma_int8_layouts():
    import probe_models as pm
    rng = np.random.default_rng(1)
    res = {}
    # --- QK^T style: A [128 x 128] K-major SW128, B [N=128 x K=128] K-major SW128 (D = 128)
    A, B = _i8(rng, 128, 128), _i8(rng, 128, 128)
    exp = A.astype(np.int32) @ B.astype(np.int32).T
    idv = pm.idesc(2, 1, 1, 0, 0, 128, 128)
    got = pm.run_mma(pm.image_rows(A.view(np.uint8), 2), pm.image_rows(B.view(np.uint8), 2), 128, idesc_v=idv, kind=1, n_mma=4)
    res["i8_kmajor_sw128"] = bool((got.numpy() == exp).all())
    # --- D = 64: rows of 64 bytes, SW64, SBO = 512
    A6, B6 = _i8(rng, 128, 64), _i8(rng, 128, 64)
    exp6 = A6.astype(np.int32) @ B6.astype(np.int32).T
    got = pm.run_mma(pm.image_rows(A6.view(np.uint8), 4), pm.image_rows(B6.view(np.uint8), 4), 128, idesc_v=idv, kind=1,
                     n_mma=2, a_sbo=512, a_layout=4, b_sbo=512, b_layout=4)
    res["i8_kmajor_sw64"] = bool((got.numpy() == exp6).all())
    # --- PV style: A = P [128 x keys=128] K-major SW128, B = V [keys=128][D=128] MN-major SW128, 4096 B per k-step
    P, V = rng.integers(0, 128, size=(128, 128), dtype=np.int8), _i8(rng, 128, 128)
    expv = P.astype(np.int32) @ V.astype(np.int32)
    idm = pm.idesc(2, 1, 1, 0, 1, 128, 128)
    for name, lbo, sbo in (("lbo16_sbo1024", 16, 1024), ("lbo1024_sbo16", 1024, 16), ("lbo4096_sbo1024", 4096, 1024),
                           ("lbo1024_sbo4096", 1024, 4096)):
        got = pm.run_mma(pm.image_rows(P.view(np.uint8), 2), pm.image_rows(V.view(np.uint8), 2), 128, idesc_v=idm, kind=1,
                         n_mma=4, b_lbo=lbo, b_sbo=sbo, b_kstep=4096)
        res["i8_B_mnmajor_sw128_" + name] = bool((got.numpy() == expv).all())
    # --- PV with D = 64: V rows of 64 bytes (SW64), N = 64; k-step = 32 keys * 64 B = 2048
    V6 = _i8(rng, 128, 64)
    expv6 = P.astype(np.int32) @ V6.astype(np.int32)
    idm6 = pm.idesc(2, 1, 1, 0, 1, 128, 64)
    for name, lbo, sbo in (("lbo16_sbo512", 16, 512), ("lbo512_sbo16", 512, 16), ("lbo2048_sbo512", 2048, 512)):
        got = pm.run_mma(pm.image_rows(P.view(np.uint8), 2), pm.image_rows(V6.view(np.uint8), 4), 64, idesc_v=idm6, kind=1,
                         n_mma=4, b_lbo=lbo, b_sbo=sbo, b_layout=4, b_kstep=2048)
        res["i8_B_mnmajor_sw64_" + name] = bool((got.numpy() == expv6).all())
    # --- transposed A (backward: dV = P^T dO): A stored [K = q rows][M = keys] (MN-major SW128), B = dO [q][D] MN-major
    Pm, dO = _i8(rng, 128, 128), _i8(rng, 128, 128)          # Pm[q][key]
    expt = Pm.astype(np.int32).T @ dO.astype(np.int32)
    idt = pm.idesc(2, 1, 1, 1, 1, 128, 128)
    for name, lbo, sbo in (("lbo16_sbo1024", 16, 1024), ("lbo4096_sbo1024", 4096, 1024)):
        got = pm.run_mma(pm.image_rows(Pm.view(np.uint8), 2), pm.image_rows(dO.view(np.uint8), 2), 128, idesc_v=idt, kind=1,
                         n_mma=4, a_lbo=lbo, a_sbo=sbo, a_kstep=4096, b_lbo=lbo, b_sbo=sbo, b_kstep=4096)
        res["i8_A_mnmajor_sw128_" + name] = bool((got.numpy() == expt).all())
    # --- TS mode: A = P from TMEM, 4 int8 per 32-bit column (little endian), 8 columns per k-step
    a_t = P.view(np.uint8).reshape(128, 32, 4).copy().view(np.uint32).reshape(128, 32)
    got = pm.run_mma(a_t, pm.image_rows(V.view(np.uint8), 2), 128, idesc_v=idm, kind=1, n_mma=4, b_kstep=4096,
                     a_in_tmem=1, a_tmem_cols=32, a_tmem_kstep_cols=8)
    res["i8_TS_packed4"] = bool((got.numpy() == expv).all())
    REPORT["i8"] = res
    _save()
    need = ["i8_kmajor_sw128", "i8_kmajor_sw64"]
    assert all(res[k] for k in need), res
    assert any(v for k, v in res.items() if k.startswith("i8_B_mnmajor_sw128")), res


def test_umma_f16_layouts():
    import probe_models as pm
    rng = np.random.default_rng(2)
    res = {}
    f16 = lambda r, c: (rng.standard_normal((r, c)) * 0.5).astype(np.float16)
    A, B = f16(128, 64), f16(128, 64)
    exp = A.astype(np.float32) @ B.astype(np.float32).T
    idv = pm.idesc(1, 0, 0, 0, 0, 128, 128)           # f32 acc, f16 x f16, K-major
    got = pm.run_mma(pm.image_rows(A.view(np.uint8), 2), pm.image_rows(B.view(np.uint8), 2), 128, idesc_v=idv, kind=0, n_mma=4)
    err = np.abs(got.numpy().view(np.float32) - exp).max()
    res["f16_kmajor_sw128"] = float(err)
    # bf16 P [128 x 128 keys] (two K atoms of 64 keys) x V [keys][D=64] bf16 MN-major (one 128 B atom along N)
    to_bf = lambda x: torch.from_numpy(x).to(torch.bfloat16)
    P = to_bf(np.abs(rng.standard_normal((128, 64))).astype(np.float32))
    V = to_bf(rng.standard_normal((64, 64)).astype(np.float32))
    expv = (P.float() @ V.float()).numpy()
    idm = pm.idesc(1, 1, 1, 0, 1, 128, 64)
    img = lambda t, lay: pm.image_rows(t.view(torch.int16).numpy().view(np.uint8).reshape(t.shape[0], -1), lay)
    for name, lbo, sbo in (("lbo16_sbo1024", 16, 1024), ("lbo2048_sbo1024", 2048, 1024)):
        got = pm.run_mma(img(P, 2), img(V, 2), 64, idesc_v=idm, kind=0, n_mma=4, b_lbo=lbo, b_sbo=sbo, b_kstep=2048)
        res["bf16_B_mnmajor_sw128_" + name] = float(np.abs(got.numpy().view(np.float32) - expv).max())
    # D = 128 bf16 V: two 128-byte atoms along N, each [keys][128 B]; LBO = atom stride = keys * 128 B
    V2 = to_bf(rng.standard_normal((64, 128)).astype(np.float32))
    expv2 = (P.float() @ V2.float()).numpy()
    v_img = np.concatenate([img(V2[:, :64].contiguous(), 2), img(V2[:, 64:].contiguous(), 2)])
    idm2 = pm.idesc(1, 1, 1, 0, 1, 128, 128)
    for name, lbo, sbo in (("lbo8192_sbo1024", 8192, 1024), ("lbo1024_sbo8192", 1024, 8192)):
        got = pm.run_mma(img(P, 2), v_img, 128, idesc_v=idm2, kind=0, n_mma=4, b_lbo=lbo, b_sbo=sbo, b_kstep=2048)
        res["bf16_B_mnmajor_2atoms_" + name] = float(np.abs(got.numpy().view(np.float32) - expv2).max())
    # TS mode bf16: 2 values per 32-bit column, 8 columns per k-step (K = 16)
    a_t = P.view(torch.int16).numpy().view(np.uint16).reshape(128, 32, 2).copy().view(np.uint32).reshape(128, 32)
    got = pm.run_mma(a_t, img(V, 2), 64, idesc_v=idm, kind=0, n_mma=4, b_kstep=2048, a_in_tmem=1, a_tmem_cols=32,
                     a_tmem_kstep_cols=8)
    res["bf16_TS_packed2"] = float(np.abs(got.numpy().view(np.float32) - expv).max())
    # transposed A in 16-bit (backward dV = P^T dO, dK = dS^T Q): A stored [K = q][M = keys] bf16, M = 128 -> two atoms
    Pq = to_bf(rng.standard_normal((64, 128)).astype(np.float32))      # [q = 64][keys = 128]
    dO = to_bf(rng.standard_normal((64, 64)).astype(np.float32))       # [q = 64][D = 64]
    expt = (Pq.float().T @ dO.float()).numpy()
    a_img = np.concatenate([img(Pq[:, :64].contiguous(), 2), img(Pq[:, 64:].contiguous(), 2)])
    idt = pm.idesc(1, 1, 1, 1, 1, 128, 64)
    for name, lbo, sbo in (("lbo8192_sbo1024", 8192, 1024), ("lbo1024_sbo8192", 1024, 8192)):
        got = pm.run_mma(a_img, img(dO, 2), 64, idesc_v=idt, kind=0, n_mma=4, a_lbo=lbo, a_sbo=sbo, a_kstep=2048,
                         b_lbo=16, b_sbo=1024, b_kstep=2048)
        res["bf16_A_mnmajor_2atoms_" + name] = float(np.abs(got.numpy().view(np.float32) - expt).max())
    REPORT["f16"] = res
    _save()
    assert res["f16_kmajor_sw128"] < 1e-2, res
