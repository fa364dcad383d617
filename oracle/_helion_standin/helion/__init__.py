"""Eager stand-in for the `helion` package -- TEST INFRASTRUCTURE ONLY.

The reference modules (/root/reference/attention_{int8,bf16,jvp}.py) do
`import helion` at top level; Helion 0.2.7 is not installable here (no network).
Their kernel bodies are ordinary PyTorch-op code over tiles, so running them
eagerly defines their semantics.  This package lets the UNMODIFIED reference
files import and run on CPU so that (a) the oracle restatements in `oracle/` can
be pinned bit-for-bit against the real source and (b) golden fixtures can be
generated (oracle/make_golden.py).  It is never imported by the product package.

Only the API the reference touches is provided:
  helion.kernel(**kw), helion.Config(block_sizes=[...]), helion.cdiv,
  helion.language.{tile,zeros,full,dot,register_tunable},
  helion.autotuner.PowerOfTwoFragment, helion._testing.{DEVICE,run_example}.
"""
from __future__ import annotations

import functools

import torch
from torch.overrides import TorchFunctionMode

from . import language as language  # noqa: F401
from . import _state


class Config:
    def __init__(self, block_sizes=None, **kw):
        self.block_sizes = list(block_sizes) if block_sizes is not None else None
        self.extra = kw


def cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


def _untile(idx):
    T = language.Tile
    if isinstance(idx, T):
        return slice(idx.begin, idx.end)
    if isinstance(idx, tuple):
        return tuple(_untile(i) for i in idx)
    return idx


class _EagerMode(TorchFunctionMode):
    """Rewrites Tile objects to slices and supplies the mixed-dtype baddbmm Helion lowers to tl.dot."""

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.Tensor.__getitem__:
            return func(args[0], _untile(args[1]))
        if func is torch.Tensor.__setitem__:
            return func(args[0], _untile(args[1]), args[2])
        if func is torch.baddbmm and not kwargs:
            acc, a, b = args
            if a.dtype != acc.dtype or b.dtype != acc.dtype:
                # tl.dot(a, b, acc): low-precision operands, fp32 accumulate
                return acc + torch.bmm(a.to(acc.dtype), b.to(acc.dtype))
        return func(*args, **kwargs)


def kernel(fn=None, *, config=None, **_kw):
    def deco(f):
        @functools.wraps(f)
        def wrapper(*args, **kwargs):
            sizes = _state.override_block_sizes
            if sizes is None and config is not None:
                sizes = config.block_sizes
            _state.begin_call(sizes)
            try:
                with _EagerMode():
                    return f(*args, **kwargs)
            finally:
                _state.end_call()

        wrapper.__wrapped_reference__ = f
        return wrapper

    return deco(fn) if fn is not None else deco
