"""helion.language stand-in (eager)."""
from __future__ import annotations

import itertools
import sys

import torch

from . import _state


class Tile:
    __slots__ = ("begin", "end", "block_size", "id")

    def __init__(self, begin, end, block_size, idx):
        self.begin, self.end, self.block_size, self.id = begin, end, block_size, idx

    def __index__(self):  # lets `hl.zeros((tile, D))`-style sizes work through int()
        return self.end - self.begin

    def __repr__(self):
        return f"Tile[{self.begin}:{self.end}]"


def _tiles_1d(n, bs):
    return [Tile(b, min(b + bs, n), bs, i) for i, b in enumerate(range(0, n, bs))]


def tile(sizes, block_size=None):
    frame = sys._getframe(1)
    site = (frame.f_code.co_filename, frame.f_lineno)
    if isinstance(sizes, (list, tuple)):
        nd = len(sizes)
        if block_size is None:
            bss = _state.sizes_for_site(site, nd)
        else:
            bss = list(block_size) if isinstance(block_size, (list, tuple)) else [block_size] * nd
        return itertools.product(*[_tiles_1d(int(n), int(b)) for n, b in zip(sizes, bss)])
    bs = block_size if block_size is not None else _state.sizes_for_site(site, 1)[0]
    return iter(_tiles_1d(int(sizes), int(bs)))


def _shape(shape):
    return [s.end - s.begin if isinstance(s, Tile) else int(s) for s in shape]


def zeros(shape, dtype=torch.float32, device=None):
    return torch.zeros(_shape(shape), dtype=dtype, device=device)


def full(shape, value, dtype=torch.float32, device=None):
    return torch.full(_shape(shape), value, dtype=dtype, device=device)


def dot(a, b, acc=None):
    if a.dtype == torch.int8 and b.dtype == torch.int8:
        # exact: |sum| <= 127*127*K < 2^24 for K <= 1040
        assert a.shape[-1] <= 1024
        r = torch.matmul(a.to(torch.float32), b.to(torch.float32)).to(torch.int32)
    else:
        r = torch.matmul(a.to(torch.float32), b.to(torch.float32))
    return r if acc is None else acc + r


def register_tunable(name, fragment):
    return int(_state.tunable_overrides.get(name, fragment.default))
