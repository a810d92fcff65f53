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
"""
from __future__ import annotations

import functools
from collections.abc import Iterable
from typing import Any, Callable, List, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import IterableDataset

from ray_shuffling_data_loader_b200.dataset import ShufflingDataset
from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.runtime.chunks import field_tensor


class TorchShufflingDataset(IterableDataset):
    """
    A PyTorch shuffling dataset that yields batches upon iteration.

    This dataset will kick off shuffling for max_concurrent_epochs epochs at
    construction time.

    Args:
        filenames (str): Paths to input Parquet files.
        num_epochs (int): Number of training epochs.
        num_trainers (int): Number of trainer workers.
        batch_size (int): Size of the batches that the iterator should yield.
        rank (int): The worker rank of the current process.
        drop_last (Optional[bool]): Whether to drop the last batch if it's
            incomplete (smaller than batch_size). Default is False.
        num_reducers (Optional[int]): The number of shuffler reducers.
        max_concurrent_epochs (Optional[int]): The maximum number of epochs
            whose shuffling stages should execute concurrently. Default is 2.
        feature_columns (List[Any]): The feature columns' names.
        feature_shapes (Optional[List[Any]]): The shape for each
            feature. If provided, it should match the size of feature_columns.
        feature_types (Optional[List[torch.dtype]]): The data type for each
            feature. If provided, it should match the size of feature_columns.
        label_column (Any): The label column name.
        label_shape (Optional[int]): The shape for the label data.
        label_type (Optional[torch.dtype]): The data type for the label data.
        packed_features (bool): yield ``(features[B, F], label)`` with one
            matrix instead of a list of ``(B, 1)`` views (needs a single
            feature dtype). Extension; default False = reference contract.
        row_align (int): opt-in padding of the packed row pitch to a multiple
            of this power of two (e.g. 128 = one L2 line); see
            ``ops/layout.py::build_layout``.
        fp8_block_scale (bool): with ``feature_types`` all
            ``torch.float8_e4m3fn``, emit MX-style block-scaled fp8 (one UE8M0
            scale per 32 features, see ``ops/fp8.py``); the batch then yields
            ``((payload, scales), label)`` in packed mode.
        seed, backend, **engine_options: see ``ShufflingDataset``.
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
                 row_align: int = 0,
                 **dataset_options):
        super().__init__()
        spec = _normalize_torch_data_spec(feature_columns, feature_shapes,
                                          feature_types, label_column,
                                          label_shape, label_type)
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

    @property
    def dataset(self) -> ShufflingDataset:
        return self._ds

    def set_epoch(self, epoch):
        """
        Set the current training epoch. This should be called before
        constructing the iterator on this dataset (e.g. before the
        enumerate(train_loader) call).

        Args:
            epoch (int) The epoch number for the training epoch that is about
                to start.
        """
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

    def __iter__(self):
        from ray_shuffling_data_loader_b200.dataset import BatchSpan
        cache = {}      # id(epoch buffer) -> (features, label) views of the WHOLE buffer
        for item in iter(self._ds):
            if isinstance(item, BatchSpan) and hasattr(item.buffer.data, "_offset"):
                item = item.packed()      # chunk shipped from another process
            if isinstance(item, BatchSpan):
                buf = item.buffer
                views = cache.get(id(buf))
                if views is None:
                    if len(cache) > 8:
                        cache.clear()
                    whole = buf.view(0, buf.rows)
                    if isinstance(whole, np.ndarray):
                        whole = torch.from_numpy(whole)
                    self._layout = buf.layout
                    views = cache[id(buf)] = (buf, packed_to_tensors(
                        whole, buf.layout, self._spec, self._packed_features))
                feats, label = views[1]
                a, b = item.start, item.stop
                # per batch: one row-slice per tensor of views built once per epoch
                if isinstance(feats, list):
                    f = [t[a:b] for t in feats]
                elif isinstance(feats, tuple):
                    f = tuple(t[a:b] for t in feats)
                else:
                    f = feats[a:b]
                yield f, (label[a:b] if label is not None else None)
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


def torch_layout(schema, spec, fp8_block_scale: bool = False, row_align: int = 0,
                 reorder: bool = False) -> L.RowLayout:
    """Row layout for a Torch data spec: features in the given order, then the
    label; each source column is cast to its requested dtype. With ``reorder``
    (tensors are looked up by column name, so storage order is free) the largest
    class of columns the TMA scatter kernel can take in one launch is stored
    first - e.g. an int64 id in front of 40 float features no longer sends the
    whole row to the generic kernel."""
    (feature_columns, feature_shapes, feature_types, label_column, label_shape,
     label_type) = spec
    cols = []
    for name, dtype in list(zip(feature_columns, feature_types)) + [(label_column, label_type)]:
        if name is None:
            continue
        if name not in schema:
            raise KeyError(f"column {name!r} not found in the Parquet schema "
                           f"{list(schema)}")
        src_code, width = schema[name]
        cols.append((name, src_code, L.code_from_torch(dtype), max(1, width)))
    if reorder and not fp8_block_scale and len({c[0] for c in cols}) == len(cols):
        cols = [cols[i] for i in L.tma_friendly_order(cols)]
    return L.build_layout(cols, fp8_block_scale=fp8_block_scale, row_align=row_align)


def packed_to_tensors(packed: torch.Tensor, layout: L.RowLayout, spec,
                      packed_features: bool = False):
    """``uint8[B, pitch]`` -> ``(features, label)`` views (no copies)."""
    (feature_columns, feature_shapes, feature_types, label_column, label_shape,
     label_type) = spec
    pitch = layout.row_pitch
    if packed_features:
        first = layout.field(feature_columns[0])
        last = layout.field(feature_columns[-1])
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
        for col, shape in zip(feature_columns, feature_shapes):
            t = field_tensor(packed, layout.field(col), pitch)
            if shape is not None:
                t = t.reshape(*(-1, *shape))
            features.append(t)
    if label_column is None:
        return features, None
    label = field_tensor(packed, layout.field(label_column), pitch)
    if label_shape:
        label = label.reshape(-1, label_shape)
    return features, label


def dataframe_to_tensor_factory(
        feature_columns: List[Any] = None,
        feature_shapes: Optional[List[Any]] = None,
        feature_types: Optional[List[torch.dtype]] = None,
        label_column: Any = None,
        label_shape: Optional[int] = None,
        label_type: Optional[torch.dtype] = None,
) -> Callable:
    """
    Returns a Pandas DataFrame --> PyTorch Tensor converter, using the
    provided data spec to do the conversion (kept for users who iterate
    ``ShufflingDataset(output="pandas")`` themselves; the fused kernel path
    never builds a DataFrame).
    """
    spec = _normalize_torch_data_spec(feature_columns, feature_shapes,
                                      feature_types, label_column, label_shape,
                                      label_type)
    return functools.partial(
        convert_to_tensor,
        feature_columns=spec[0],
        feature_shapes=spec[1],
        feature_types=spec[2],
        label_column=spec[3],
        label_shape=spec[4],
        label_type=spec[5])


def _normalize_torch_data_spec(
        feature_columns: List[Any] = None,
        feature_shapes: Optional[List[Any]] = None,
        feature_types: Optional[List[torch.dtype]] = None,
        label_column: Any = None,
        label_shape: Optional[int] = None,
        label_type: Optional[torch.dtype] = None):
    """
    Normalize the provided Torch data spec, returning sensible defaults for
    unspecified parameters (same rules as reference torch_dataset.py:144-201).
    """
    # Convert to list for convenience.
    if not isinstance(feature_columns, list):
        feature_columns = [feature_columns]

    if feature_shapes:
        if not isinstance(feature_shapes, list):
            feature_shapes = [feature_shapes]
        assert len(feature_columns) == len(feature_shapes), \
            "The feature_shapes size must match the feature_columns"
        feature_shapes = [
            s if (s is None or isinstance(s, Iterable)) else [s]
            for s in feature_shapes]
    else:
        feature_shapes = [None] * len(feature_columns)

    if feature_types:
        if not isinstance(feature_types, list):
            feature_types = [feature_types]
        assert len(feature_columns) == len(feature_types), \
            "The feature_types size must match the feature_columns"
        assert all(isinstance(dtype, torch.dtype) for dtype in feature_types), \
            "All value in feature_types should be torch.dtype instance"
    else:
        feature_types = [torch.float] * len(feature_columns)

    if not label_type:
        label_type = torch.float

    return (feature_columns, feature_shapes, feature_types, label_column,
            label_shape, label_type)


def convert_to_tensor(df, feature_columns: List[Any],
                      feature_shapes: List[Any],
                      feature_types: List[torch.dtype], label_column: Any,
                      label_shape: Optional[int], label_type: torch.dtype):
    """Host-side DataFrame -> tensors conversion (reference semantics, fixed
    for numpy >= 1.24 where ``np.object`` no longer exists)."""
    feature_tensor = []
    for col, shape, dtype in zip(feature_columns, feature_shapes,
                                 feature_types):
        column = df[col].values
        if column.dtype == object:
            if isinstance(column[0], np.ndarray):
                column = np.stack(column)
            elif isinstance(column[0], (list, tuple)):
                column = list(column)
            else:
                raise Exception(
                    f"Column {col}'s type: {type(column[0])} is not supported."
                    " It must be numpy built in type or numpy object of "
                    "(ndarray, list, tuple)")
        t = torch.as_tensor(column, dtype=dtype)
        if shape is not None:
            t = t.view(*(-1, *shape))
        else:
            t = t.view(-1, 1)
        feature_tensor.append(t)

    label_df = df[label_column].values
    label_tensor = torch.as_tensor(label_df, dtype=label_type)
    if label_shape:
        label_tensor = label_tensor.view(-1, label_shape)
    else:
        label_tensor = label_tensor.view(-1, 1)
    return feature_tensor, label_tensor


def _smoke_main():
    """``python -m ray_shuffling_data_loader_b200.torch_dataset``: the
    reference's smoke driver (``torch_dataset.py:239-309``)."""
    import shutil
    import tempfile
    from ray_shuffling_data_loader_b200.stats import human_readable_size
    from ray_shuffling_data_loader_b200.data_generation import (generate_data,
                                                                DATA_SPEC)
    num_rows = 10**6
    num_files = 10
    data_dir = tempfile.mkdtemp()
    filenames, num_bytes = generate_data(num_rows, num_files, 1, 0.0, data_dir)
    print(f"Generated {len(filenames)} files containing {num_rows} rows, "
          f"totalling {human_readable_size(num_bytes)}.")
    num_epochs, num_trainers, batch_size, rank = 4, 1, 20000, 0
    num_reducers, max_concurrent_epochs = 8, 2
    feature_columns = list(DATA_SPEC.keys())
    numpy_to_torch_dtype = {
        np.bool_: torch.bool, np.uint8: torch.uint8, np.int8: torch.int8,
        np.int16: torch.int16, np.int32: torch.int32, np.int64: torch.int64,
        np.float16: torch.float16, np.float32: torch.float32,
        np.float64: torch.float64}
    feature_types = [numpy_to_torch_dtype[dtype] for _, _, dtype in DATA_SPEC.values()]
    label_column = feature_columns.pop()
    label_type = feature_types.pop()
    print(f"Creating Torch shuffling dataset with {batch_size} batch size, "
          f"{num_epochs} epochs, {num_reducers} reducers, and {num_trainers} "
          "trainers.")
    print(f"Should consume {num_rows // batch_size} batches.")
    ds = TorchShufflingDataset(
        filenames, num_epochs, num_trainers, batch_size, rank,
        num_reducers=num_reducers, max_concurrent_epochs=max_concurrent_epochs,
        feature_columns=feature_columns, feature_types=feature_types,
        label_column=label_column, label_type=label_type)
    for epoch in range(num_epochs):
        ds.set_epoch(epoch)
        for batch_idx, (data, targets) in enumerate(ds):
            print(f"Epoch {epoch} - consuming batch {batch_idx}: "
                  f"{len(data)} features, {len(targets)} samples")
    print("Done consuming batches.")
    shutil.rmtree(data_dir)


if __name__ == "__main__":
    _smoke_main()
