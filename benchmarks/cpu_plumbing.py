#!/usr/bin/env python
"""BASELINE.json config 1 - "ShufflingDataset num_trainers=1 num_reducers=2 on
CPU, 4 synthetic Parquet files" - for both implementations, no GPU involved.

    python benchmarks/cpu_plumbing.py --impl ours        # backend="cpu" (C++ host runtime)
    python benchmarks/cpu_plumbing.py --impl ours-numpy  # same, pure numpy golden path
    python benchmarks/cpu_plumbing.py --impl reference   # unmodified reference on baseline/ray_shim

Same files (DATA_SPEC schema, written by our generator), same iteration: every
batch of every epoch is pulled through the public ``ShufflingDataset`` iterator as
a pandas DataFrame and its ``key`` column is summed (exactly-once check). The
clock covers construction (the reference starts shuffling there) to the last
batch; rows/s = num_epochs * num_rows / seconds, the reference's own definition
(``stats.py:396-397``). One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "ours-numpy", "reference"], default="ours")
    ap.add_argument("--num-rows", type=int, default=10**6)          # the reference's smoke size
    ap.add_argument("--num-files", type=int, default=4)
    ap.add_argument("--num-epochs", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=20_000)
    ap.add_argument("--num-reducers", type=int, default=2)
    ap.add_argument("--data-dir", default=os.path.join(tempfile.gettempdir(), "rsdl_cpu_plumbing"))
    a = ap.parse_args()

    sys.path.insert(0, ROOT)
    from ray_shuffling_data_loader_b200.data_generation import generate_data
    marker = os.path.join(a.data_dir, f".ok_{a.num_rows}_{a.num_files}")
    if not os.path.exists(marker):
        files, _ = generate_data(a.num_rows, a.num_files, 2, 0.0, a.data_dir, seed=11)
        open(marker, "w").close()
    files = [os.path.join(a.data_dir, f"input_data_{i}.parquet.snappy") for i in range(a.num_files)]

    if a.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "ray_shim"))
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import ray
        from ray_shuffling_data_loader import ShufflingDataset
        ray.init(num_cpus=os.cpu_count())
        kwargs = {}
    else:
        from ray_shuffling_data_loader_b200 import ShufflingDataset
        if a.impl == "ours-numpy":
            os.environ["RSDL_CPU_NATIVE"] = "0"
        kwargs = dict(backend="cpu", seed=1)

    t0 = time.perf_counter()
    ds = ShufflingDataset(files, a.num_epochs, 1, a.batch_size, 0,
                          num_reducers=a.num_reducers, max_concurrent_epochs=2, **kwargs)
    rows = 0
    key_sum = 0
    for epoch in range(a.num_epochs):
        ds.set_epoch(epoch)
        for df in ds:
            rows += len(df)
            key_sum += int(df["key"].sum())
    seconds = time.perf_counter() - t0
    want = a.num_epochs * a.num_rows
    print(json.dumps({
        "impl": a.impl, "config": "ShufflingDataset num_trainers=1 num_reducers=%d, %d files, CPU"
        % (a.num_reducers, a.num_files), "num_rows": a.num_rows, "num_epochs": a.num_epochs,
        "batch_size": a.batch_size, "seconds": seconds, "rows_per_sec": rows / seconds,
        "rows_delivered": rows, "rows_expected": want,
        "exactly_once": key_sum == a.num_epochs * (a.num_rows * (a.num_rows - 1) // 2),
        "cpus": os.cpu_count()}), flush=True)
    if a.impl == "reference":
        try:
            ray.shutdown()
        finally:
            os._exit(0)


if __name__ == "__main__":
    main()
