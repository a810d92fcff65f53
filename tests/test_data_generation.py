"""Synthetic Parquet generator (C14) - with the reference's quirks fixed."""
import os

import pytest

import numpy as np
import pandas as pd
import pyarrow.parquet as pq

from ray_shuffling_data_loader_b200.data_generation import (DATA_SPEC, float_spec,
                                                            generate_data)
from ray_shuffling_data_loader_b200.runtime import ingest


def test_exact_file_count_unique_keys_and_row_groups(tmp_path):
    # 1003 % 4 != 0: the reference would write 5 files here.
    filenames, nbytes = generate_data(1003, 4, 3, 0.0, str(tmp_path), seed=0)
    assert len(filenames) == 4
    assert [os.path.basename(f) for f in filenames] == [
        f"input_data_{i}.parquet.snappy" for i in range(4)]
    df = pd.concat([pd.read_parquet(f) for f in filenames])
    assert len(df) == 1003
    assert np.array_equal(df["key"].to_numpy(), np.arange(1003))   # global, unique
    assert list(df.columns) == ["key"] + list(DATA_SPEC.keys())
    for col, (low, high, dtype) in DATA_SPEC.items():
        assert df[col].dtype == np.dtype(dtype)
        assert df[col].min() >= low and df[col].max() < max(high, low + 1)
    assert nbytes == 1003 * 21 * 8
    md = pq.ParquetFile(filenames[0]).metadata
    assert md.num_row_groups == 3
    assert md.row_group(0).column(0).compression == "SNAPPY"


def test_float_spec_honours_dtype_and_seed(tmp_path):
    spec = float_spec(8, np.float32)
    a, _ = generate_data(500, 2, 1, 0.0, str(tmp_path / "a"), data_spec=spec, seed=5)
    b, _ = generate_data(500, 2, 1, 0.0, str(tmp_path / "b"), data_spec=spec, seed=5)
    da, db = pd.read_parquet(a[0]), pd.read_parquet(b[0])
    assert all(da[c].dtype == np.float32 for c in spec)
    assert da.equals(db)
    assert list(da.columns) == ["key"] + [f"f{i}" for i in range(7)] + ["labels"]


def test_ingest_row_ranges(tmp_path):
    filenames, _ = generate_data(1000, 3, 2, 0.0, str(tmp_path), seed=1)
    index = ingest.scan_files(filenames)
    assert index.num_rows == 1000 and len(index.row_groups) == 6
    assert index.file_rows == [334, 333, 333]
    full = pd.concat([pd.read_parquet(f) for f in filenames])
    # a range that starts and ends inside row groups, with column projection
    t = ingest.load_table(index, 150, 777, columns=["key", "labels"], num_threads=3)
    assert t.num_rows == 627 and t.global_offset == 150
    assert np.array_equal(t.columns["key"], np.arange(150, 777))
    assert np.array_equal(t.columns["labels"], full["labels"].to_numpy()[150:777])
    empty = ingest.load_table(index, 10, 10, columns=["key"])
    assert empty.num_rows == 0 and len(empty.columns["key"]) == 0


def test_row_group_skew(tmp_path):
    """max_row_group_skew (asserted to be 0.0 upstream, data_generation.py:15):
    sizes vary within +-skew of the mean, totals and keys are untouched."""
    import pyarrow.parquet as pq
    from ray_shuffling_data_loader_b200.data_generation import generate_data
    files, _ = generate_data(40_000, 2, 8, 0.5, str(tmp_path), seed=3)
    keys = []
    for f in files:
        md = pq.ParquetFile(f).metadata
        sizes = [md.row_group(i).num_rows for i in range(md.num_row_groups)]
        assert len(sizes) == 8 and sum(sizes) == 20_000
        assert len(set(sizes)) > 1
        mean = 20_000 / 8
        assert min(sizes) >= mean * 0.5 / 1.5 - 1 and max(sizes) <= mean * 1.5 / 0.5 + 1
        keys.append(pq.read_table(f, columns=["key"]).column("key").to_numpy())
    assert np.array_equal(np.concatenate(keys), np.arange(40_000))
    with pytest.raises(ValueError):
        generate_data(100, 1, 2, 1.5, str(tmp_path / "bad"))
