#!/usr/bin/env python
"""Cold path: Parquet -> pinned host -> HBM for bench.py's table on
one GPU, per decode-thread count: total ingest seconds, its pinned-allocation and decode
parts, rows/s. ``decode_seconds`` overlaps the H2D copies (every decoded row-group slice is
handed to the copy engine immediately); what remains after it is the copy tail.

    python tools/ingest_bench.py [--threads 8 16 32 64]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--batch-size", type=int, default=250_000)
    ap.add_argument("--schema", default="f32")
    ap.add_argument("--data-dir", default=os.environ.get("RSDL_BENCH_DIR", "/tmp/rsdl_bench"))
    ap.add_argument("--threads", type=int, nargs="+", default=[8, 16, 32, 64])
    a = ap.parse_args()
    import bench
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    torch.cuda.set_device(0)
    bench.generate_my_share(a, 0, 1)
    _, files = bench.dataset_files(a, 1)
    cols = [f"f{i}" for i in range(a.cols - 1)] + ["labels"]

    def layout_fn(schema):
        return L.build_layout([(c, schema[c][0], L.DT_F32, 1) for c in cols])
    plan_args = dict(num_trainers=1, num_reducers=1, batch_size=a.batch_size, drop_last=False)
    for nt in a.threads:
        for rep in range(2):
            eng = DeviceShuffleEngine(files, plan_args, layout_fn, 1, num_threads=nt)
            eng._ensure_ingested(0)
            print(json.dumps({"num_threads": nt, "rep": rep, "cpus": len(os.sched_getaffinity(0)),
                              "ingest_seconds": eng.ingest_seconds,
                              "pinned_alloc_seconds": eng.pinned_alloc_seconds,
                              "decode_seconds": eng.decode_seconds,
                              "rows_per_sec": a.rows_per_gpu / eng.ingest_seconds,
                              "gb_per_sec": a.rows_per_gpu * a.cols * 4 / eng.ingest_seconds / 1e9}),
                  flush=True)
            eng.close()


if __name__ == "__main__":
    main()
