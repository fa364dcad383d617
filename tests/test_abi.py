"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/qattn.h declares,
the ctypes table covers the header, argument validation fails loudly, and the product package has no CPU path and
never imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols(name="qattn.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from quantizedattention_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/qattn.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib().qa_version() >= 100
    # the product library carries no probe / debug entry points, its sources no environment switches
    assert not [s for s in syms if s.startswith(("qa_probe", "qa_debug"))]
    for s in _header_symbols("qattn_dev.h"):
        assert not hasattr(L, s), f"{s} is a development symbol but libqattn.so exports it"
    csrc = os.path.join(ROOT, "quantizedattention_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f"{f}: kernel selection must be an argument, not an env var"


def test_dev_library_exports_dev_header():
    from quantizedattention_b200 import _lib, build
    build.build(dev=True)
    L = ctypes.CDLL(_lib.DEV_LIB_PATH)
    dev = _header_symbols("qattn_dev.h")
    assert set(dev) == set(_lib.DEV_SIGNATURES), "ctypes dev table and qattn_dev.h disagree"
    for s in dev + _header_symbols():
        assert hasattr(L, s), s


def test_argument_validation_needs_no_gpu():
    from quantizedattention_b200 import _lib
    L = _lib.lib()
    z = ctypes.c_void_p(0)
    assert L.qa_quant_block(z, z, z, z, 128, 96, 32, 128, 0, z) == -1         # D not in {64,128}
    assert b"D must be" in L.qa_last_error()
    assert L.qa_quant_block(z, z, z, z, 100, 64, 32, 128, 0, z) == -1         # rows not a multiple of blk
    assert L.qa_quant_block(z, z, z, z, 128, 64, 32, 128, 3, z) == -1         # rounding in {0, 1, 2}
    assert L.qa_fp8_fwd(*([z] * 9), 1, 100, 128, 128, z) == -1                # fp8 forward: S % 128
    assert L.qa_int8_fwd(*([z] * 12), 1, 128, 128, 128, 128, 16, 1, 0, z) == -1  # Bkv in {32,64,128,256}
    assert L.qa_int8_bwd(*([z] * 14), 1, 128, 128, 32, 128, 0, z) == -1       # Bq = Bkv = 128
    assert L.qa_bf16_fwd(*([z] * 5), 1, 100, 128, 128, 0, 1, z) == -1         # S % 128
    assert L.qa_bf16_fwd_ex(*([z] * 5), 1, 128, 128, 128, 0, 0, -1.0, z) == -1  # rescale_tau outside [0, 16]
    assert L.qa_bf16_fwd_ex(*([z] * 5), 1, 128, 128, 128, 0, 0, 32.0, z) == -1
    assert L.qa_jvp_fwd(*([z] * 9), 1, 128, 128, 96, 1, z) == -1              # D in {64,128}
    assert L.qa_bf16_fwd_ragged(*([z] * 5), 1, 128, 256, 100, 128, 0, 0, 8.0, z) == -1   # Sk_valid in (Sk - 128, Sk]
    assert L.qa_jvp_fwd_ragged(*([z] * 9), 1, 128, 128, 0, 128, 1, z) == -1
    assert L.qa_bf16_bwd_ragged(*([z] * 10), 1, 256, 300, 128, 0, 0, z) == -1


def test_no_cpu_fallback_and_no_oracle_in_product():
    from quantizedattention_b200 import attention_bf16, attention_int8, attention_jvp
    x = torch.randn(1, 1, 128, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        attention_int8.helion_atten_int8_hl_dot_fwd(x.half(), x.half(), x.half())
    with pytest.raises(RuntimeError, match="no CPU path"):
        attention_bf16.helion_atten_bf16_fwd_training(x.half(), x.half(), x.bfloat16(), False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        attention_jvp.helion_attention_jvp_forward_fp32(x, x, x, x, x, x)
    with pytest.raises(TypeError):
        attention_bf16.helion_atten_bf16_fwd_training(x.half(), x.half(), x.half(), False)   # LEDGER B-11
    pkg = os.path.join(ROOT, "quantizedattention_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} must not use the oracle"


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from quantizedattention_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libqattn.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_bench_reference_arm_prints_contract_json():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a one-head sample: one JSON line with
    the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-heads", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_torch_library_ops_register_without_gpu():
    """the operator schemas and fake kernels exist on import (shape inference needs no device)."""
    import torch
    from quantizedattention_b200 import torch_ops  # noqa: F401
    assert "Tensor q, Tensor k, Tensor v" in str(torch.ops.qattn.sage_int8_fwd.default._schema)
    with torch._subclasses.FakeTensorMode():
        q = torch.empty((2, 4, 256, 128), dtype=torch.float16, device="cuda")
        out = torch.ops.qattn.sage_int8_fwd(q, q, q, 128, 128, False)
        assert out[0].shape == (2, 4, 256, 128) and out[3].dtype == torch.int8 and out[6].shape == (2 * 4 * 256 // 128,)
        O, lse = torch.ops.qattn.flash_bf16_fwd(q, q, q.to(torch.bfloat16), True)
        assert O.dtype == torch.float32 and lse.shape == (8, 256)


def test_tools_and_entry_points_compile():
    """Every script that is only ever run on the GPU box at least parses here."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]:
        py_compile.compile(f, doraise=True)
