"""TorchShufflingDataset: ``(List[Tensor], Tensor)`` batches (component C10).

Same constructor, ``set_epoch`` and output contract as the reference
(``ray_shuffling_data_loader/torch_dataset.py:14-236``): one tensor per feature
column shaped ``(B, 1)`` (or ``(B, *feature_shape)``), a label tensor shaped
``(B, 1)`` (or ``(B, label_shape)``), default dtype ``torch.float``.

The reference converts every batch on the host, column by column, with
``torch.as_tensor(column, dtype)`` (``torch_dataset.py:204-236``) and the
training loop then copies each tensor to the GPU from pageable memory
(``examples/horovod/ray_torch_shuffle.py:204-207``). Here the dtype cast and the
row packing happen inside the shuffle kernel's epilogue, so a batch is a
``[B, row_pitch]`` byte matrix already in HBM and every feature tensor is a
strided zero-copy view of it; ``packed_features=True`` additionally exposes the
single ``[B, F]`` matrix most models want.

The data spec (which columns, their shapes and dtypes) is a small value object,
``TensorSpec``; the packed-row layout is derived from it (``torch_layout``) and
the same object drives both the device path (``packed_to_tensors``: views) and
the host-side helper kept for API parity (``convert_to_tensor``).
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import IterableDataset

from ray_shuffling_data_loader_b200.dataset import ShufflingDataset
from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.runtime.chunks import field_tensor


@dataclass(frozen=True)
class ColumnSpec:
    """One output tensor: source column, trailing shape (``None`` = ``(1,)``), dtype."""
    name: Any
    shape: Optional[Tuple[int, ...]]
    dtype: torch.dtype

    def trailing(self) -> Tuple[int, ...]:
        return (1,) if self.shape is None else self.shape


@dataclass(frozen=True)
class TensorSpec:
    """The feature/label contract of ``TorchShufflingDataset`` in one place.

    Built by ``TensorSpec.build`` from the reference's six keyword arguments
    (reference ``torch_dataset.py:144-201`` for the defaulting rules: scalars are
    promoted to one-element lists, missing shapes mean ``(B, 1)``, missing dtypes
    mean ``torch.float``); malformed specs raise ``ValueError``/``TypeError``
    (they are user errors, and ``assert`` disappears under ``python -O``)."""
    features: Tuple[ColumnSpec, ...]
    label: Optional[ColumnSpec]

    @staticmethod
    def build(feature_columns=None, feature_shapes=None, feature_types=None,
              label_column=None, label_shape=None, label_type=None) -> "TensorSpec":
        names = list(feature_columns) if isinstance(feature_columns, (list, tuple)) \
            else [feature_columns]
        shapes = _broadcast_arg(feature_shapes, len(names), "feature_shapes")
        dtypes = _broadcast_arg(feature_types, len(names), "feature_types")
        feats = []
        for name, shape, dtype in zip(names, shapes, dtypes):
            if dtype is None:
                dtype = torch.float
            elif not isinstance(dtype, torch.dtype):
                raise TypeError(f"feature_types entries must be torch.dtype, got {dtype!r}")
            feats.append(ColumnSpec(name, _as_shape(shape), dtype))
        label = None
        if label_column is not None:
            if label_type is not None and not isinstance(label_type, torch.dtype):
                raise TypeError(f"label_type must be a torch.dtype, got {label_type!r}")
            label = ColumnSpec(label_column, _as_shape(label_shape or None),
                               label_type or torch.float)
        return TensorSpec(tuple(feats), label)

    @property
    def columns(self) -> List[ColumnSpec]:
        return list(self.features) + ([self.label] if self.label is not None else [])


def _broadcast_arg(value, n: int, what: str) -> list:
    """``None``/empty -> ``[None] * n``; a scalar -> one-element list; a list must
    have one entry per feature column."""
    if value is None or (isinstance(value, (list, tuple)) and len(value) == 0):
        return [None] * n
    # a bare shape tuple such as (3, 32, 32) for a single feature is one entry
    if not isinstance(value, list):
        value = [value]
    if len(value) != n:
        raise ValueError(f"{what} has {len(value)} entries for {n} feature columns")
    return list(value)


def _as_shape(shape) -> Optional[Tuple[int, ...]]:
    if shape is None:
        return None
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(d) for d in shape)


class TorchShufflingDataset(IterableDataset):
    """Per-epoch globally shuffled ``(features, label)`` tensor batches.

    Positional arguments and their order are the reference's
    (``torch_dataset.py:45-59``) so existing call sites keep working:

    ``filenames``              input Parquet files (their concatenation is the table)
    ``num_epochs``             epochs that will be iterated
    ``num_trainers``           data-parallel consumers; each gets a disjoint 1/T of every epoch
    ``batch_size``             rows per yielded batch (the last one may be short)
    ``rank``                   which trainer this process is
    ``drop_last``              skip the short final batch
    ``num_reducers``           destination chunks per epoch with their own completion flag
    ``max_concurrent_epochs``  epoch ring depth: how many shuffles may be in flight
    ``feature_columns`` / ``feature_shapes`` / ``feature_types``
                               one output tensor per column: ``(B, 1)`` or ``(B, *shape)``,
                               dtype default ``torch.float`` (cast inside the shuffle kernel)
    ``label_column`` / ``label_shape`` / ``label_type``
                               the label tensor, ``(B, 1)`` or ``(B, label_shape)``

    The first ``max_concurrent_epochs`` shuffles start in the constructor.

    Keyword-only extensions:

    ``packed_features``  yield one ``[B, F]`` matrix instead of F ``(B, 1)`` views
                         (needs a single feature dtype)
    ``row_align``        pad the packed row pitch to a multiple of this power of two
                         (128 = one L2 line; see ``ops/layout.py::build_layout``). Default
                         ``None`` = auto: only rows of 96..127 bytes are padded (to 128);
                         ``0`` = never pad
    ``fp8_block_scale``  with all-``float8_e4m3fn`` features: MX-style block scaling (one
                         UE8M0 scale per 32 features, ``ops/fp8.py``); packed mode then
                         yields ``((payload, scales), label)``
    ``seed``, ``backend``, engine options: see ``ShufflingDataset``.
    """

    def __init__(self,
                 filenames: List[str],
                 num_epochs: int,
                 num_trainers: int,
                 batch_size: int,
                 rank: int,
                 drop_last: bool = False,
                 num_reducers=None,
                 max_concurrent_epochs=2,
                 feature_columns: List[Any] = None,
                 feature_shapes: Optional[List[Any]] = None,
                 feature_types: Optional[List[torch.dtype]] = None,
                 label_column: Any = None,
                 label_shape: Optional[int] = None,
                 label_type: Optional[torch.dtype] = None,
                 *,
                 packed_features: bool = False,
                 fp8_block_scale: bool = False,
                 row_align: Optional[int] = None,
                 **dataset_options):
        super().__init__()
        spec = TensorSpec.build(feature_columns, feature_shapes, feature_types,
                                label_column, label_shape, label_type)
        self._spec = spec
        self._packed_features = packed_features
        self._fp8_block_scale = fp8_block_scale
        self._layout_fn = functools.partial(torch_layout, spec=spec,
                                            fp8_block_scale=fp8_block_scale,
                                            row_align=row_align,
                                            reorder=not packed_features)
        self._ds = ShufflingDataset(
            filenames,
            num_epochs,
            num_trainers,
            batch_size,
            rank,
            drop_last=drop_last,
            num_reducers=num_reducers,
            max_concurrent_epochs=max_concurrent_epochs,
            output="span",
            layout_fn=self._layout_fn,
            **dataset_options)
        self._transform: Optional[Callable] = None
        self._layout = None
        self._slot_views = {}

    @property
    def dataset(self) -> ShufflingDataset:
        return self._ds

    def set_epoch(self, epoch):
        """Select the epoch the next ``iter()`` will read. Mandatory before every
        pass and the value must change between passes (``ValueError`` otherwise),
        exactly like the reference (``torch_dataset.py:78-88``)."""
        self._ds.set_epoch(epoch)

    def state_dict(self):
        return self._ds.state_dict()

    def load_state_dict(self, state):
        self._ds.load_state_dict(state)

    def _get_transform(self):
        if self._transform is None:
            eng = self._ds.engine
            layout = eng.layout if eng is not None else None
            self._transform = layout
        return self._transform

    def _buffer_views(self, buf):
        """-> ``((features, label) views of the WHOLE epoch buffer, {(start, stop): batch})``.

        On the GPU engine an epoch buffer is a slot of the device epoch ring: the same
        memory every ``max_concurrent_epochs`` epochs, so both the whole-buffer views
        and the per-batch slices are kept across epochs (keyed by device pointer and
        row count) and a steady-state step costs two dictionary look-ups instead of
        building tensor views - the Python side of a step matters once a step is only
        tens of microseconds of device time (profiles/README.md, round 2)."""
        whole = buf.view(0, buf.rows)
        if isinstance(whole, np.ndarray):
            whole = torch.from_numpy(whole)
        self._layout = buf.layout
        key = None
        if whole.is_cuda:
            key = (whole.data_ptr(), buf.rows, whole.shape[-1])
            hit = self._slot_views.get(key)
            if hit is not None:
                return hit
        entry = (packed_to_tensors(whole, buf.layout, self._spec, self._packed_features), {})
        if key is not None:
            if len(self._slot_views) >= 16:          # a ring never has that many slots
                self._slot_views.clear()
            self._slot_views[key] = entry
        return entry

    def __iter__(self):
        from ray_shuffling_data_loader_b200.dataset import BatchSpan
        cache = {}      # id(epoch buffer) -> per-buffer entry (see _buffer_views)
        for item in iter(self._ds):
            if isinstance(item, BatchSpan) and hasattr(item.buffer.data, "_offset"):
                # chunk shipped from the owning process (this rank only connected
                # to its queue): the pickled chunk carries the layout with it
                self._layout = item.buffer.layout
                item = item.packed()
            if isinstance(item, BatchSpan):
                buf = item.buffer
                entry = cache.get(id(buf))
                if entry is None:
                    if len(cache) > 8:
                        cache.clear()
                    entry = cache[id(buf)] = (buf, self._buffer_views(buf))
                (feats, label), batches = entry[1]
                key = (item.start, item.stop)
                got = batches.get(key)
                if got is None:
                    a, b = key
                    # one row-slice per tensor of views built once per buffer
                    if isinstance(feats, list):
                        f = [t[a:b] for t in feats]
                    elif isinstance(feats, tuple):
                        f = tuple(t[a:b] for t in feats)
                    else:
                        f = feats[a:b]
                    got = batches[key] = (f, (label[a:b] if label is not None else None))
                f, lab = got
                yield (list(f) if isinstance(f, list) else f), lab
                continue
            packed = item           # a batch that straddled two buffers (copy)
            if self._layout is None:
                eng = self._ds.engine
                if eng is None:
                    raise RuntimeError(
                        "TorchShufflingDataset needs the layout of the owning "
                        "process; construct it with rank 0 or in distributed mode")
                self._layout = eng.layout
            if isinstance(packed, np.ndarray):
                packed = torch.from_numpy(packed)
            yield packed_to_tensors(packed, self._layout, self._spec,
                                    self._packed_features)


def torch_layout(schema, spec: TensorSpec, fp8_block_scale: bool = False,
                 row_align: Optional[int] = None, reorder: bool = False) -> L.RowLayout:
    """Row layout for a Torch data spec: features in the given order, then the
    label; each source column is cast to its requested dtype. With ``reorder``
    (tensors are looked up by column name, so storage order is free) the largest
    class of columns the TMA scatter kernel can take in one launch is stored
    first - e.g. an int64 id in front of 40 float features no longer sends the
    whole row to the generic kernel."""
    cols = []
    for c in spec.columns:
        if c.name not in schema:
            raise KeyError(f"column {c.name!r} not found in the Parquet schema "
                           f"{list(schema)}")
        src_code, width = schema[c.name]
        cols.append((c.name, src_code, L.code_from_torch(c.dtype), max(1, width)))
    if reorder and not fp8_block_scale and len({c[0] for c in cols}) == len(cols):
        cols = [cols[i] for i in L.tma_friendly_order(cols)]
    lay = L.build_layout(cols, fp8_block_scale=fp8_block_scale, row_align=row_align or 0)
    if row_align is None and 96 <= lay.row_pitch < 128:
        # auto: a row just short of a 128-byte line (the reference's own DATA_SPEC as
        # torch.float features: 21 x 4 B -> 96 B) is padded to the full line. Measured
        # (profiles/README.md round 2): scatter 0.80 -> 0.74 ms in local HBM and
        # 2.65 -> 2.30 ms over NVLink for 12.5 M rows although 33 % more bytes move -
        # every row then is exactly one line / one NVLink write packet. Wider or much
        # narrower rows gain nothing or lose; ``row_align=0`` keeps the tight pitch.
        lay = L.build_layout(cols, fp8_block_scale=fp8_block_scale, row_align=128)
    return lay


def packed_to_tensors(packed: torch.Tensor, layout: L.RowLayout, spec: TensorSpec,
                      packed_features: bool = False):
    """``uint8[B, pitch]`` -> ``(features, label)`` views (no copies)."""
    pitch = layout.row_pitch
    if packed_features:
        first = layout.field(spec.features[0].name)
        last = layout.field(spec.features[-1].name)
        code = first.dst_code
        nelem = (last.offset + last.dst_bytes - first.offset) // L.itemsize(code)
        whole = L.Field("features", first.src_code, code, first.offset, nelem)
        features = field_tensor(packed, whole, pitch)
        if layout.scale_offset >= 0:
            nblk = (nelem + 31) // 32
            scales = packed[:, layout.scale_offset:layout.scale_offset + nblk]
            features = (features, scales)
    else:
        features = []
        for c in spec.features:
            t = field_tensor(packed, layout.field(c.name), pitch)
            if c.shape is not None:
                t = t.reshape(-1, *c.shape)
            features.append(t)
    if spec.label is None:
        return features, None
    label = field_tensor(packed, layout.field(spec.label.name), pitch)
    if spec.label.shape is not None:
        label = label.reshape(-1, *spec.label.shape)
    return features, label


# ---------------------------------------------------------------------------
# host-side DataFrame -> tensors (API parity with the reference's helpers)
# ---------------------------------------------------------------------------

def _cells_to_array(values: np.ndarray, column) -> np.ndarray:
    """Column values -> one dense ndarray. Object columns (what pandas gives for
    Parquet list columns: one ndarray / list / tuple per row) are stacked."""
    if values.dtype != object:
        return values
    if len(values) == 0:
        return np.empty((0,), dtype=np.float32)
    probe = values[0]
    if not isinstance(probe, (np.ndarray, list, tuple)):
        raise TypeError(f"column {column!r} holds {type(probe).__name__} cells; only numeric "
                        "columns and ndarray/list/tuple cells of equal length are supported")
    return np.stack([np.asarray(v) for v in values])


def _to_tensor(df, c: ColumnSpec) -> torch.Tensor:
    arr = _cells_to_array(df[c.name].to_numpy(), c.name)
    if not (arr.flags.c_contiguous and arr.flags.writeable):
        arr = np.array(arr, order="C")
    t = torch.from_numpy(arr).to(c.dtype)
    # same result shapes as the device path (``packed_to_tensors``): a scalar
    # column is (B, 1), a list column without an explicit shape is (B, width)
    return t.reshape(len(arr), -1) if c.shape is None else t.reshape(len(arr), *c.shape)


def convert_to_tensor(df, spec=None, *more, **spec_kwargs):
    """``DataFrame -> (features: List[Tensor], label: Tensor)`` on the host, for
    callers that iterate ``ShufflingDataset(output="pandas")`` themselves (role of
    reference ``torch_dataset.py:204-236``). Accepts a prebuilt ``TensorSpec``, the
    six spec fields as keyword arguments, or - like the reference's signature -
    positionally (``feature_columns, feature_shapes, feature_types, label_column,
    label_shape, label_type``). The GPU path never comes through here."""
    if not isinstance(spec, TensorSpec):
        positional = (() if spec is None else (spec,)) + more
        spec = TensorSpec.build(*positional, **spec_kwargs)
    features = [_to_tensor(df, c) for c in spec.features]
    label = _to_tensor(df, spec.label) if spec.label is not None else None
    return features, label


def dataframe_to_tensor_factory(feature_columns=None, feature_shapes=None, feature_types=None,
                                label_column=None, label_shape=None,
                                label_type=None) -> Callable:
    """A ``DataFrame -> tensors`` converter bound to one data spec (reference
    ``torch_dataset.py:95-141``); the spec is validated once, here."""
    spec = TensorSpec.build(feature_columns, feature_shapes, feature_types,
                            label_column, label_shape, label_type)
    return functools.partial(convert_to_tensor, spec=spec)


def _smoke_main(argv: Optional[Sequence[str]] = None) -> int:
    """``python -m ray_shuffling_data_loader_b200.torch_dataset``: end-to-end smoke
    run on generated ``DATA_SPEC`` files (role of the reference's ``__main__``
    block, ``torch_dataset.py:239-309``) - but it *checks* what it consumes: row
    count per epoch, tensor shapes/dtypes, and that two epochs differ."""
    import argparse
    import tempfile
    from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC, generate_data
    ap = argparse.ArgumentParser(description=_smoke_main.__doc__)
    ap.add_argument("--num-rows", type=int, default=10**6)
    ap.add_argument("--num-files", type=int, default=10)
    ap.add_argument("--num-epochs", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=20000)
    ap.add_argument("--num-reducers", type=int, default=8)
    ap.add_argument("--backend", default=None, choices=[None, "cpu", "cuda"])
    a = ap.parse_args(argv)
    to_torch = {np.dtype(np.int64): torch.int64, np.dtype(np.float64): torch.float64,
                np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}
    *feature_columns, label_column = DATA_SPEC
    dtypes = [to_torch[np.dtype(DATA_SPEC[c][2])] for c in DATA_SPEC]
    with tempfile.TemporaryDirectory() as data_dir:
        files, nbytes = generate_data(a.num_rows, a.num_files, 1, 0.0, data_dir)
        print(f"{len(files)} files, {a.num_rows} rows, {nbytes / 1e6:.1f} MB decoded; "
              f"{a.num_epochs} epochs of {-(-a.num_rows // a.batch_size)} batches")
        ds = TorchShufflingDataset(
            files, a.num_epochs, 1, a.batch_size, 0, num_reducers=a.num_reducers,
            feature_columns=feature_columns, feature_types=dtypes[:-1],
            label_column=label_column, label_type=dtypes[-1], backend=a.backend)
        first_of_epoch = []
        for epoch in range(a.num_epochs):
            ds.set_epoch(epoch)
            rows = 0
            for i, (features, label) in enumerate(ds):
                if i == 0:
                    first_of_epoch.append(features[0][:8, 0].clone())
                if len(features) != len(feature_columns) or label.shape[1] != 1:
                    raise AssertionError("tensor contract violated")
                rows += label.shape[0]
            if rows != a.num_rows:
                raise AssertionError(f"epoch {epoch}: {rows} rows, expected {a.num_rows}")
            print(f"epoch {epoch}: {rows} rows in {i + 1} batches")
        if a.num_epochs > 1 and all(torch.equal(first_of_epoch[0], t) for t in first_of_epoch[1:]):
            raise AssertionError("every epoch started with the same rows: not shuffled")
    print("ok")
    return 0


if __name__ == "__main__":
    raise SystemExit(_smoke_main())
