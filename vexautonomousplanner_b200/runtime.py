"""Process-wide default engine used by the drop-in mirrors of the reference classes (B = 1 wrappers)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .engine import Engine, _p

_engine: Optional[Engine] = None


def get_engine() -> Engine:
    """The default Engine on cuda:0 (dt = 0.01, dd = 0.005, 1000-sample tables).  Raises without a GPU."""
    global _engine
    if _engine is None:
        _engine = Engine("cuda:0")
    return _engine


def set_engine(engine: Optional[Engine]) -> None:
    global _engine
    _engine = engine


def dev(a, dtype=torch.float64) -> torch.Tensor:
    eng = get_engine()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(eng.device)


class Keep:
    """Keeps temporaries alive until a kernel that reads them has been enqueued (the caching allocator may hand a
    freed block to the very next allocation, i.e. before the launch, if the Python object is already gone)."""

    def __init__(self):
        self.items = []

    def __call__(self, a, dtype=torch.float64):
        t = a if isinstance(a, torch.Tensor) else dev(a, dtype)
        self.items.append(t)
        return _p(t)


def stream():
    return get_engine()._stream()


def check(rc, what=""):
    _lib.check(rc, what)


__all__ = ["get_engine", "set_engine", "dev", "stream", "check", "C", "_p", "Keep"]
