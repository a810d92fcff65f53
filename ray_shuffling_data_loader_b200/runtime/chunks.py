"""Epoch buffers, reducer-chunk handles and packed-row views.

In the reference a reducer's output is a pandas DataFrame living in Ray's
plasma store and the queue carries its ``ObjectRef`` (reference
``shuffle.py:171-200``, ``dataset.py:133-139``). Here a reducer chunk is a row
range of one trainer's *epoch buffer* - a ``[rows, row_pitch]`` byte matrix that
the shuffle kernel fills in place (HBM in GPU mode, numpy on the CPU backend) -
and the queue carries ``ShuffledChunk`` handles onto it. ``wait()`` is the
analogue of ``ray.wait(..., fetch_local=True)`` + ``ray.get``: it blocks until
the producers' completion flags for the chunk have fired.
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ray_shuffling_data_loader_b200.ops import layout as L


class EpochBuffer:
    """One trainer's shuffled rows for one epoch.

    ``data`` is ``uint8[rows, pitch]`` (numpy array or CUDA torch tensor).
    ``ready`` callables block until the given chunk's data has landed."""

    def __init__(self, epoch: int, trainer: int, rows: int, layout: L.RowLayout,
                 data: Any, device: str = "cpu",
                 wait_fn: Optional[Callable[[Optional[float]], None]] = None,
                 release_fn: Optional[Callable[[], None]] = None):
        self.epoch = epoch
        self.trainer = trainer
        self.rows = rows
        self.layout = layout
        self.data = data
        self.device = device
        self._wait_fn = wait_fn
        self._release_fn = release_fn
        self._ready = threading.Event()
        self._rows_cv = threading.Condition()
        self._rows_ready = 0          # rows [0, _rows_ready) have landed (CPU engine, K7)
        self._error: Optional[BaseException] = None
        self._released = False
        if wait_fn is None and device != "cpu":
            self._ready.set()

    # producer side (CPU engine) -----------------------------------------
    def mark_ready(self, error: Optional[BaseException] = None):
        self._error = error
        with self._rows_cv:
            if error is None:
                self._rows_ready = self.rows
            self._ready.set()
            self._rows_cv.notify_all()

    def mark_rows_ready(self, row_stop: int):
        """Rows ``[0, row_stop)`` are final (the host engine fills a buffer chunk by
        chunk, in order): wakes consumers that wait for a prefix only."""
        with self._rows_cv:
            self._rows_ready = max(self._rows_ready, int(row_stop))
            self._rows_cv.notify_all()

    # consumer side ---------------------------------------------------------
    def wait(self, timeout: Optional[float] = None, row_stop: Optional[int] = None):
        """Block (or, in stream mode, make the current CUDA stream wait) until rows
        ``[0, row_stop)`` - default: the whole buffer - have landed."""
        if self._wait_fn is not None:
            if row_stop is None:
                self._wait_fn(timeout)
            else:
                self._wait_fn(timeout, row_stop)
            return
        need = self.rows if row_stop is None else min(int(row_stop), self.rows)
        with self._rows_cv:
            ok = self._rows_cv.wait_for(
                lambda: self._rows_ready >= need or self._ready.is_set(), timeout)
        if not ok:
            raise TimeoutError(
                f"epoch {self.epoch} buffer of trainer {self.trainer} not ready "
                f"after {timeout}s")
        if self._error is not None:
            raise self._error

    def release(self):
        """The trainer is done with every row of this epoch (device analogue of
        the reference's ``task_done`` + ``queue.join()``, batch_queue.py:413-414)."""
        if not self._released:
            self._released = True
            if self._release_fn is not None:
                self._release_fn()

    def view(self, start: int, stop: int):
        return self.data[start:stop]


class ShuffledChunk:
    """Handle to rows ``[row_start, row_stop)`` of an ``EpochBuffer``."""

    def __init__(self, buffer: EpochBuffer, chunk_index: int, row_start: int,
                 row_stop: int):
        self.buffer = buffer
        self.chunk_index = chunk_index
        self.row_start = row_start
        self.row_stop = row_stop

    @property
    def epoch(self) -> int:
        return self.buffer.epoch

    @property
    def trainer(self) -> int:
        return self.buffer.trainer

    @property
    def layout(self) -> L.RowLayout:
        return self.buffer.layout

    def __len__(self) -> int:
        return self.row_stop - self.row_start

    def wait(self, timeout: Optional[float] = None) -> "ShuffledChunk":
        """Wait for THIS chunk only: the completion flags of the destination-chunk
        pass that delivers ``[.., row_stop)`` (reference analogue: ``ray.wait`` on
        one reducer output, ``dataset.py:133-139``)."""
        self.buffer.wait(timeout, self.row_stop)
        return self

    def packed(self):
        """``uint8[rows, pitch]`` view (blocks until ready)."""
        self.wait()
        return self.buffer.view(self.row_start, self.row_stop)

    def to_pandas(self):
        return packed_to_dataframe(_to_numpy(self.packed()), self.layout)

    # Chunks cross process boundaries only on the CPU backend (a remote
    # ``BatchQueue`` consumer); ship the rows, not the handle.
    def __reduce__(self):
        data = _to_numpy(self.packed()).copy()
        return (_rebuild_chunk, (self.epoch, self.trainer, self.chunk_index,
                                 self.row_start, self.row_stop, self.layout, data))


def _rebuild_chunk(epoch, trainer, chunk_index, row_start, row_stop, layout, data):
    buf = EpochBuffer(epoch, trainer, row_stop, layout, None)
    # Only this chunk's rows travelled: index relative to row_start.
    buf.data = _OffsetRows(data, row_start)
    buf.mark_ready()
    return ShuffledChunk(buf, chunk_index, row_start, row_stop)


class _OffsetRows:
    def __init__(self, data, offset):
        self._data = data
        self._offset = offset

    def __getitem__(self, sl):
        return self._data[sl.start - self._offset:sl.stop - self._offset]


def _to_numpy(packed) -> np.ndarray:
    if isinstance(packed, np.ndarray):
        return packed
    return packed.detach().cpu().numpy()


# ---------------------------------------------------------------------------
# packed rows -> user-facing objects
# ---------------------------------------------------------------------------

_FIELD_DESC = np.dtype([("ptr", "<u8"), ("src_code", "<u4"), ("dst_code", "<u4"),
                        ("dst_off", "<u4"), ("width", "<u4")])
_NATIVE = {"C": None, "pool": None, "tried": False}


def _native_unpacker():
    """(extension module, shared worker pool) for ``host_unpack_fields``, or (None, None)
    when the extension is not built / disabled (``RSDL_CPU_NATIVE=0``)."""
    if not _NATIVE["tried"]:
        _NATIVE["tried"] = True
        import os
        if os.environ.get("RSDL_CPU_NATIVE", "1") != "0":
            try:
                from ray_shuffling_data_loader_b200 import _C
                _NATIVE["C"] = _C
                _NATIVE["pool"] = _C.HostPool(max(1, min(16, (os.cpu_count() or 2))))
            except Exception:
                _NATIVE["C"] = None
    return _NATIVE["C"], _NATIVE["pool"]


def unpack_columns(packed: np.ndarray, layout: L.RowLayout) -> Dict[str, np.ndarray]:
    """Packed rows -> one owned, contiguous array per display field (``[n]`` or
    ``[n, width]`` in the field's storage dtype). With the extension built this is one
    multi-threaded transposition pass (``host_unpack_fields``); otherwise numpy copies."""
    fields = list(layout.display_fields)
    n = int(packed.shape[0])
    C, pool = _native_unpacker()
    if C is None or n < 4096 or not packed.flags.c_contiguous:
        return {f.name: L.unpack_field(packed, f) for f in fields}
    out, desc = {}, np.zeros(len(fields), dtype=_FIELD_DESC)
    for i, f in enumerate(fields):
        dt = L.numpy_storage_dtype(f.dst_code)
        arr = np.empty((n,) if f.width == 1 else (n, f.width), dtype=dt)
        out[f.name] = arr
        desc[i] = (arr.ctypes.data, f.src_code, f.dst_code, f.offset, f.width)
    C.host_unpack_fields(pool, packed.ctypes.data, n, int(packed.shape[1]), desc.ctypes.data,
                         len(fields))
    return out


def packed_to_dataframe(packed: np.ndarray, layout: L.RowLayout):
    """Rebuild a pandas DataFrame (what the reference's iterator yields,
    ``dataset.py:108-188``) from packed rows. List-valued fields become object
    columns of ndarrays, like Parquet list columns read by pandas. The columns are
    owned copies (the packed buffer may be recycled) handed to pandas without a
    second, consolidating copy."""
    import pandas as pd
    cols = unpack_columns(packed, layout)
    for f in layout.display_fields:
        if f.width > 1:
            vals = cols[f.name]
            obj = np.empty(len(vals), dtype=object)
            for i in range(len(vals)):
                obj[i] = vals[i]
            cols[f.name] = obj
    return pd.DataFrame(cols, copy=False)


def field_tensor(packed, f: L.Field, pitch: int):
    """Zero-copy strided view ``[B, width]`` of one field of a packed batch
    (torch tensor in, torch tensor out; works on CPU and CUDA)."""
    import torch
    size = L.itemsize(f.dst_code)
    dt = L.torch_dtype(f.dst_code)
    rows = packed.shape[0]
    if pitch % size == 0 and f.offset % size == 0:
        typed = packed.view(torch.uint8).view(dt) if size > 1 else packed.view(dt)
        # typed: [B, pitch/size]
        start = f.offset // size
        return typed[:, start:start + f.width]
    # Misaligned (never produced by build_layout): fall back to a copy.
    raw = packed[:, f.offset:f.offset + f.dst_bytes].contiguous()
    return raw.view(dt).view(rows, f.width)


class DeviceBatch:
    """A batch of packed rows plus its layout: what ``ShufflingDataset`` yields
    in GPU mode instead of a pandas DataFrame (rows never leave HBM).

    ``batch[name]`` -> strided tensor view ``[B]`` (or ``[B, width]``),
    ``len(batch)``, ``batch.columns``, ``batch.to_pandas()`` (explicit D2H)."""

    def __init__(self, packed, layout: L.RowLayout):
        self.packed = packed
        self.layout = layout

    def __len__(self):
        return int(self.packed.shape[0])

    @property
    def columns(self) -> List[str]:
        return self.layout.names

    def __getitem__(self, name: str):
        f = self.layout.field(name)
        t = field_tensor(self.packed, f, self.layout.row_pitch)
        return t[:, 0] if f.width == 1 else t

    def to_pandas(self):
        return packed_to_dataframe(_to_numpy(self.packed), self.layout)

    def matrix(self):
        """``[B, F]`` view when every field shares one dtype (else ``None``)."""
        code = self.layout.uniform_code
        if code is None:
            return None
        import torch
        size = L.itemsize(code)
        nelem = self.layout.payload_bytes // size
        typed = self.packed.view(L.torch_dtype(code)) if size > 1 else self.packed.view(L.torch_dtype(code))
        return typed[:, :nelem]
