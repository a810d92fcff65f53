"""Parquet ingest (kernel-table row K10): files -> columnar host table.

The reference re-reads every file every epoch with ``pd.read_parquet(filename)``
inside each mapper task (reference ``shuffle.py:151``; "loads data from disk
once per epoch", ``shuffle.py:46-48``), one core per file, all columns.

Here ingest happens once per dataset: each process decodes only the row range it
owns (``ShufflePlan.source_range``) with column projection, row-group parallel
on a thread pool (pyarrow's C++ decoder releases the GIL), straight into
preallocated column buffers - pinned host memory when a GPU runtime is supplied
so the later ``cudaMemcpyAsync`` is a true async DMA. The decoded table is then
kept resident (HBM or pinned host), so epochs >= 1 never touch Parquet again.
"""
from __future__ import annotations

import threading
import timeit
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from ray_shuffling_data_loader_b200.ops import layout as L


@dataclass
class RowGroupRef:
    file_index: int
    filename: str
    row_group: int
    num_rows: int
    global_start: int   # global index of the row group's first row


@dataclass
class DatasetIndex:
    """Footer-only scan of the input files (cheap; every rank does it)."""
    filenames: List[str]
    row_groups: List[RowGroupRef]
    num_rows: int
    schema: Dict[str, Tuple[int, int]]      # name -> (dtype code, width)
    file_rows: List[int] = field(default_factory=list)


def _arrow_field_spec(f: pa.Field) -> Tuple[int, int]:
    t = f.type
    if pa.types.is_fixed_size_list(t):
        return L.code_from_numpy(t.value_type.to_pandas_dtype()), t.list_size
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        # variable-size list *type* (what pandas writes for ndarray-valued cells):
        # the per-row length is resolved from the footer by ``scan_files`` and must
        # be constant, like the reference's np.stack requires (torch_dataset.py:212-221)
        return L.code_from_numpy(t.value_type.to_pandas_dtype()), -1
    return L.code_from_numpy(t.to_pandas_dtype()), 1


def _list_width_from_footer(md, name: str, filename: str) -> int:
    """Per-row element count of a ``list<item>`` column, from the first non-empty
    row group's column-chunk statistics (leaf values / rows). The layout (row
    pitch, kernel plan) is built from the schema before any data is decoded, so
    the width has to be known here; ``load_table`` re-validates it on the data."""
    for rg in range(md.num_row_groups):
        g = md.row_group(rg)
        if g.num_rows == 0:
            continue
        for c in range(g.num_columns):
            col = g.column(c)
            if col.path_in_schema.split(".")[0] != name:
                continue
            if col.num_values % g.num_rows:
                raise ValueError(
                    f"{filename}: list column {name!r} has {col.num_values} values in "
                    f"{g.num_rows} rows - rows must all have the same length")
            return max(1, col.num_values // g.num_rows)
    return 1


def scan_files(filenames: Sequence[str]) -> DatasetIndex:
    """Read Parquet footers: row-group sizes, global offsets and the schema.
    Global row index := position in the concatenation of ``filenames``."""
    filenames = list(filenames)
    groups: List[RowGroupRef] = []
    schema: Optional[Dict[str, Tuple[int, int]]] = None
    start = 0
    file_rows = []

    def _meta(fn):
        pf = pq.ParquetFile(fn)
        return pf.metadata, pf.schema_arrow

    with ThreadPoolExecutor(max_workers=min(16, max(1, len(filenames)))) as ex:
        metas = list(ex.map(_meta, filenames))
    for fi, (fn, (md, sch)) in enumerate(zip(filenames, metas)):
        if schema is None:
            schema = {}
            for f in sch:
                if f.name.startswith("__index_level_"):
                    continue
                code, width = _arrow_field_spec(f)
                if width < 0:
                    width = _list_width_from_footer(md, f.name, fn)
                schema[f.name] = (code, width)
        nrows = 0
        for rg in range(md.num_row_groups):
            n = md.row_group(rg).num_rows
            groups.append(RowGroupRef(fi, fn, rg, n, start + nrows))
            nrows += n
        file_rows.append(nrows)
        start += nrows
    return DatasetIndex(filenames, groups, start, schema or {}, file_rows)


@dataclass
class HostTable:
    """Columnar rows ``[global_offset, global_offset + num_rows)``."""
    columns: Dict[str, np.ndarray]      # 1-D, or [N, width] for list columns
    schema: Dict[str, Tuple[int, int]]  # name -> (code, width)
    num_rows: int
    global_offset: int
    read_durations: List[float] = field(default_factory=list)
    pinned: bool = False

    def nbytes(self) -> int:
        return int(sum(c.nbytes for c in self.columns.values()))


def _column_to_numpy(col: pa.ChunkedArray) -> np.ndarray:
    t = col.type
    if pa.types.is_fixed_size_list(t) or pa.types.is_list(t) or pa.types.is_large_list(t):
        arr = col.combine_chunks()
        flat = arr.flatten().to_numpy(zero_copy_only=False)
        n = len(arr)
        if n == 0:
            return flat.reshape(0, max(1, getattr(t, "list_size", 1)))
        if arr.null_count or len(flat) % n != 0:
            raise ValueError("list-valued columns must have a constant length per row "
                             "and no null rows")
        if not pa.types.is_fixed_size_list(t):
            lens = arr.value_lengths().to_numpy(zero_copy_only=False)
            if len(lens) and (lens.min() != lens.max()):
                raise ValueError("list-valued columns must have a constant length per row "
                                 f"(found {int(lens.min())}..{int(lens.max())})")
        return flat.reshape(n, len(flat) // n)
    if col.null_count:
        raise ValueError("null values are not supported by the shuffling loader")
    return col.to_numpy()


def load_table(index: DatasetIndex, row_start: int, row_stop: int,
               columns: Optional[Sequence[str]] = None,
               num_threads: int = 8,
               alloc: Optional[Callable[[Tuple[int, ...], np.dtype], np.ndarray]] = None,
               on_read: Optional[Callable[[float, float], None]] = None,
               copy_fn: Optional[Callable[[np.ndarray, np.ndarray], None]] = None,
               on_slice: Optional[Callable[[str, int, int, np.ndarray], None]] = None,
               prealloc: Optional[Dict[str, np.ndarray]] = None) -> HostTable:
    """Decode global rows ``[row_start, row_stop)``.

    ``alloc(shape, dtype)`` supplies the destination buffers (pinned host
    memory from the native runtime in GPU mode); default is plain numpy.
    ``on_read(total_duration, read_duration)`` is invoked per row group (feeds
    the map-stage stats, the analogue of reference ``shuffle.py:147-167``).
    ``copy_fn(dst, src)`` copies a decoded contiguous column slice into its
    (pinned) destination - the native runtime's GIL-free parallel memcpy.
    ``on_slice(name, first_row, rows, column_buffer)`` fires (on the decode thread)
    as soon as rows ``[first_row, first_row + rows)`` of a column are in place - the
    GPU engine enqueues that slice's H2D copy right there, so staging into HBM
    overlaps the decode of the remaining row groups instead of following it.
    ``prealloc`` supplies ready-made destination arrays per column (the disk-streaming
    mode's pinned staging slots) instead of calling ``alloc``."""
    names = list(columns) if columns is not None else list(index.schema.keys())
    for n in names:
        if n not in index.schema:
            raise KeyError(f"column {n!r} not in dataset schema {list(index.schema)}")
    n_local = max(0, row_stop - row_start)
    todo = [g for g in index.row_groups
            if g.global_start < row_stop and g.global_start + g.num_rows > row_start
            and g.num_rows > 0]
    schema = {n: index.schema[n] for n in names}
    bufs: Dict[str, np.ndarray] = dict(prealloc) if prealloc else {}
    alloc = alloc or (lambda shape, dt: np.empty(shape, dtype=dt))

    alloc_lock = threading.Lock()

    def _ensure(name: str, width: int, dtype) -> np.ndarray:
        with alloc_lock:
            if name not in bufs:
                shape = (n_local,) if width == 1 else (n_local, width)
                bufs[name] = alloc(shape, np.dtype(dtype))
            return bufs[name]

    # Pre-allocate fixed-width columns so decode threads can fill in place.
    for n in names:
        code, width = schema[n]
        if width >= 1:
            _ensure(n, width, L.numpy_storage_dtype(code))

    files: Dict[str, pq.ParquetFile] = {}
    durations: List[float] = []

    def _read(g: RowGroupRef):
        t0 = timeit.default_timer()
        pf = pq.ParquetFile(g.filename)
        tbl = pf.read_row_group(g.row_group, columns=names, use_threads=False)
        t1 = timeit.default_timer()
        lo = max(row_start, g.global_start) - g.global_start
        hi = min(row_stop, g.global_start + g.num_rows) - g.global_start
        dst = g.global_start + lo - row_start
        for n in names:
            arr = _column_to_numpy(tbl.column(n))
            code, width = schema[n]
            if width < 0:
                width = arr.shape[1]
                schema[n] = (code, width)
            elif arr.ndim == 2 and arr.shape[1] != width:
                raise ValueError(f"column {n}: rows of {arr.shape[1]} elements in "
                                 f"{g.filename}, the dataset's first row group had {width}")
            buf = _ensure(n, width, arr.dtype)
            part = arr[lo:hi]
            if copy_fn is not None and part.flags.c_contiguous and part.nbytes >= (1 << 20):
                copy_fn(buf[dst:dst + (hi - lo)], part)
            else:
                buf[dst:dst + (hi - lo)] = part
            if on_slice is not None and hi > lo:
                on_slice(n, dst, hi - lo, buf)
        t2 = timeit.default_timer()
        return t2 - t0, t1 - t0

    if todo:
        with ThreadPoolExecutor(max_workers=max(1, min(num_threads, len(todo)))) as ex:
            for total, read in ex.map(_read, todo):
                durations.append(read)
                if on_read is not None:
                    on_read(total, read)
    del files
    for n in names:                      # empty range / variable-width never seen
        code, width = schema[n]
        if n not in bufs:
            width = max(width, 1)
            schema[n] = (code, width)
            _ensure(n, width, L.numpy_storage_dtype(code))
    return HostTable(bufs, schema, n_local, row_start, durations)
