"""Tracing: Chrome-trace JSON spans + NVTX ranges.

The reference only has ``timeit`` brackets reported to a stats actor
(reference ``shuffle.py:147-167,190-199``). Here every engine stage can be
wrapped in ``span(name)``: it pushes an NVTX range (visible in Nsight) and, when
``RSDL_TRACE=/path/trace.json`` is set, appends a Chrome-trace "X" event that
``chrome://tracing`` / Perfetto can open.
"""
from __future__ import annotations

import atexit
import contextlib
import json
import os
import threading
import time

_EVENTS = []
_LOCK = threading.Lock()
_PATH = os.environ.get("RSDL_TRACE")
_T0 = time.perf_counter()


def enabled() -> bool:
    return _PATH is not None


def _nvtx():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.nvtx
    except Exception:
        pass
    return None


@contextlib.contextmanager
def span(name: str, **args):
    nv = _nvtx()
    if nv is not None:
        nv.range_push(name)
    t0 = time.perf_counter()
    try:
        yield
    finally:
        t1 = time.perf_counter()
        if nv is not None:
            nv.range_pop()
        if _PATH is not None:
            with _LOCK:
                _EVENTS.append({"name": name, "ph": "X", "pid": os.getpid(),
                                "tid": threading.get_ident() % 100000,
                                "ts": (t0 - _T0) * 1e6, "dur": (t1 - t0) * 1e6,
                                "args": args})


def instant(name: str, **args):
    if _PATH is not None:
        with _LOCK:
            _EVENTS.append({"name": name, "ph": "i", "s": "p", "pid": os.getpid(),
                            "tid": threading.get_ident() % 100000,
                            "ts": (time.perf_counter() - _T0) * 1e6, "args": args})


def dump(path: str = None):
    path = path or _PATH
    if path is None:
        return
    with _LOCK:
        events = list(_EVENTS)
    rank = os.environ.get("RANK")
    if rank is not None and "%r" not in path:
        root, ext = os.path.splitext(path)
        path = f"{root}.rank{rank}{ext}"
    with open(path.replace("%r", rank or "0"), "w") as f:
        json.dump({"traceEvents": events, "displayTimeUnit": "ms"}, f)


atexit.register(dump)
