"""Per-kernel device timing for the roofline report: CUDA events recorded on the launching stream around tagged
C-ABI calls (``with profiling.tag("decode_l2"): ...``).  Inactive unless a ``KernelTimer`` is installed."""
from __future__ import annotations

import contextlib
from collections import defaultdict

import torch

from . import _lib


@contextlib.contextmanager
def tag(name: str):
    prev = _lib._tag
    _lib._tag = name
    try:
        yield
    finally:
        _lib._tag = prev


class KernelTimer:
    """Collects (start, end) event pairs per tag; ``summary()`` synchronises and returns {tag: (count, total_ms)}."""

    def __init__(self, tags):
        self.tags = set(tags)
        self.events = defaultdict(list)

    @contextlib.contextmanager
    def _sink(self, tag_name, entry):
        if tag_name not in self.tags:
            yield
            return
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.events[tag_name].append((s, e))

    def __enter__(self):
        _lib._tag_sink = self._sink
        return self

    def __exit__(self, *exc):
        _lib._tag_sink = None

    def reset(self):
        self.events.clear()

    def summary(self):
        torch.cuda.synchronize()
        return {t: (len(ev), sum(s.elapsed_time(e) for s, e in ev)) for t, ev in self.events.items()}


# ---- NVTX stage ranges (profiling hooks) -----------------------------------------------------------
try:
    from torch.cuda import nvtx as _nvtx
    _nvtx.range_push("gnb.probe")
    _nvtx.range_pop()
except Exception:  # no NVTX library in this build: ranges become no-ops
    _nvtx = None


def nvtx_push(name: str) -> None:
    if _nvtx is not None:
        _nvtx.range_push(name)


def nvtx_pop() -> None:
    if _nvtx is not None:
        _nvtx.range_pop()
