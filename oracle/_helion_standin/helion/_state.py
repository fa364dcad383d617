"""Per-call state of the eager stand-in: block sizes by hl.tile call site, tunable overrides."""
from __future__ import annotations

override_block_sizes = None      # e.g. [1, 128, 128] to force tile sizes for the next calls
default_block_size = 32          # used when neither block_size= nor a config entry exists
tunable_overrides: dict = {}     # {"Bq": 128, "Bkv": 128}

_queue: list = []
_by_site: dict = {}
_stack: list = []


def begin_call(sizes):
    global _queue, _by_site
    _stack.append((_queue, _by_site))
    _queue = list(sizes) if sizes is not None else []
    _by_site = {}


def end_call():
    global _queue, _by_site
    _queue, _by_site = _stack.pop()


def sizes_for_site(site, ndim):
    if site not in _by_site:
        out = []
        for _ in range(ndim):
            out.append(_queue.pop(0) if _queue else default_block_size)
        _by_site[site] = out
    return _by_site[site]
