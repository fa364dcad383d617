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


def test_umma_layout_cases():
    """Each case runs in a child process: a wrong descriptor hypothesis can fault the CUDA context."""
    import subprocess
    import sys
    import probe_cases
    names = list(probe_cases.CASES)
    res = {}
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_cases.py")
    todo = list(names)
    while todo:
        p = subprocess.run([sys.executable, script] + todo, capture_output=True, text=True, timeout=300)
        done = []
        for line in p.stdout.splitlines():
            if line.startswith("PROBE "):
                d = json.loads(line[6:])
                res[d["case"]] = d["result"]
                done.append(d["case"])
        todo = [n for n in todo if n not in done]
        if todo and p.returncode != 0:
            res[todo[0]] = "crash: " + p.stderr.strip().splitlines()[-1][:200] if p.stderr.strip() else "crash"
            todo = todo[1:]
        elif todo:
            break
    REPORT["umma"] = res
    _save()
    print(json.dumps(res, indent=1))
    ok = lambda v: (v is True) or (isinstance(v, float) and v < 2e-2)
    for must in ("i8_kmajor_sw128", "i8_kmajor_sw64", "f16_kmajor_sw128", "nvf4_blockscaled_ss_sw64", "nvf4_blockscaled_ts"):
        assert ok(res.get(must)), res
    assert any(ok(v) for k, v in res.items() if k.startswith("i8_Bmn_sw128")), res
    assert any(ok(v) for k, v in res.items() if k.startswith("bf16_Bmn_2atoms")), res
