"""ctypes loader for the C-ABI library (include/qattn.h).  There is NO fallback: if libqattn.so is
missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqattn.so")
DEV_LIB_PATH = os.path.join(_HERE, "libqattn_dev.so")
_lib = None
_dev = None

c_void_p, c_int, c_size_t, c_ll, c_float, c_uint = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t,
                                                    ctypes.c_longlong, ctypes.c_float, ctypes.c_uint)

# name -> (restype, argtypes); must list every symbol include/qattn.h declares
SIGNATURES = {
    "qa_version": (c_int, []),
    "qa_last_error": (ctypes.c_char_p, []),
    "qa_k_mean_workspace_bytes": (c_size_t, [c_int] * 4),
    "qa_workspace_bytes": (c_size_t, [c_int] * 5),
    "qa_k_mean": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "qa_quant_block": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "qa_k_token_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "qa_int8_fwd_state": (c_int, [c_void_p] * 15 + [c_int] * 8 + [c_void_p]),
    "qa_int8_fwd": (c_int, [c_void_p] * 12 + [c_int] * 8 + [c_void_p]),
    "qa_int8_fwd_ragged": (c_int, [c_void_p] * 15 + [c_int] * 9 + [c_void_p]),
    "qa_int8_bwd_ragged": (c_int, [c_void_p] * 14 + [c_int] * 7 + [c_void_p]),
    "qa_int8_bwd": (c_int, [c_void_p] * 14 + [c_int] * 6 + [c_void_p]),
    "qa_int8_bwd_sage": (c_int, [c_void_p] * 14 + [c_int] * 6 + [c_void_p]),
    "qa_fp8_fwd": (c_int, [c_void_p] * 9 + [c_int] * 4 + [c_void_p]),
    "qa_fp4_quant_rows": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p]),
    "qa_fp4_quant_rows_ragged": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p]),
    "qa_fp4_fwd_ragged": (c_int, [c_void_p] * 11 + [c_int] * 7 + [c_float, c_void_p]),
    "qa_fp4_quant_vt": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "qa_fp4_fwd": (c_int, [c_void_p] * 11 + [c_int] * 6 + [c_void_p]),
    "qa_int8_bwd_finalize": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "qa_bwd_delta": (c_int, [c_void_p] * 4 + [c_ll, c_int, c_int, c_void_p]),
    "qa_cast_f32": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "qa_bf16_fwd": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "qa_bf16_fwd_ex": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_float, c_void_p]),
    "qa_bf16_fwd_ragged": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_float, c_void_p]),
    "qa_jvp_fwd": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_void_p]),
    "qa_jvp_fwd_ragged": (c_int, [c_void_p] * 9 + [c_int] * 6 + [c_void_p]),
    "qa_bf16_bwd": (c_int, [c_void_p] * 10 + [c_int] * 4 + [c_void_p]),
    "qa_bf16_bwd_ex": (c_int, [c_void_p] * 10 + [c_int] * 5 + [c_void_p]),
    "qa_bf16_bwd_ragged": (c_int, [c_void_p] * 10 + [c_int] * 6 + [c_void_p]),
}


# development library only (include/qattn_dev.h): hardware probes and kernel timeline hooks
DEV_SIGNATURES = {
    "qa_debug_set_int8_fwd_timeline": (c_int, [c_void_p]),
    "qa_debug_set_int8_bwd_timeline": (c_int, [c_void_p]),
    "qa_debug_set_bf16_bwd_timeline": (c_int, [c_void_p]),
    "qa_debug_set_bf16_fwd_timeline": (c_int, [c_void_p]),
    "qa_probe_tmem_bw": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "qa_probe_tmem_bw_ex": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "qa_probe_mma": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p] + [c_int] * 8 + [c_uint] + [c_int] * 6 + [c_void_p]),
    "qa_probe_mma_bs": (c_int, [c_void_p, c_int] * 4 + [c_void_p] + [c_int] * 8 + [c_uint] + [c_int] * 8 + [c_void_p]),
    "qa_probe_tma": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
}


def _load(path, tables):
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -m quantizedattention_b200.build{' --dev' if path.endswith('_dev.so') else ''}` "
            "(there is no CPU / PyTorch fallback)")
    L = ctypes.CDLL(path)
    for table in tables:
        for name, (res, args) in table.items():
            fn = getattr(L, name)          # AttributeError = header/library mismatch: fail loudly
            fn.restype, fn.argtypes = res, args
    return L


def lib():
    """The product library.  QA_DEV_LIB=1 in the environment makes every call go through the development library
    instead (same kernels compiled with the timeline hooks; tools/timeline*.py)."""
    global _lib
    if _lib is None:
        if os.environ.get("QA_DEV_LIB") == "1":
            _lib = dev_lib()
        else:
            _lib = _load(LIB_PATH, [SIGNATURES])
    return _lib


def dev_lib():
    global _dev
    if _dev is None:
        _dev = _load(DEV_LIB_PATH, [SIGNATURES, DEV_SIGNATURES])
    return _dev


def check(rc: int, what: str, L=None):
    if rc != 0:
        msg = (L or lib()).qa_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def cur_stream():
    """Raw handle of the calling thread's current CUDA stream on the current device (torch.cuda.current_stream() builds a
    Stream object through several Python layers: ~15 us per call, five calls per attention call)."""
    import torch
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
