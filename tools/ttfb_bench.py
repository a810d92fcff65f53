#!/usr/bin/env python
"""Time to first batch (K7) on one GPU, through the engine.

For the headline table shape (``bench.py``'s files: 12.5 M rows x 64 float32) and
8 reducer chunks per trainer, measures per ``chunk_passes`` setting:

* ``first_chunk_ms``  device time from the start of an epoch's shuffle to the moment
                      chunk 0's completion flags have fired (= first batch consumable)
* ``epoch_ms``        device time of the whole epoch's shuffle
* cold start          wall time from engine construction (Parquet decode + H2D ingest)
                      to chunk 0 of epoch 0 being consumable
* host streaming      ``resident="host"``: the table is re-streamed over PCIe every epoch,
                      so the first chunk completes with the last source chunk

    python tools/ttfb_bench.py [--rows 12500000 --cols 64]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--batch-size", type=int, default=250_000)
    ap.add_argument("--schema", default="f32")
    ap.add_argument("--data-dir", default=os.environ.get("RSDL_BENCH_DIR", "/tmp/rsdl_bench"))
    ap.add_argument("--reducers", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=6)
    a = ap.parse_args()
    import bench
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.runtime.chunks import ShuffledChunk
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    torch.cuda.set_device(0)
    bench.generate_my_share(a, 0, 1)
    _, files = bench.dataset_files(a, 1)
    cols = [f"f{i}" for i in range(a.cols - 1)] + ["labels"]

    def layout_fn(schema):
        return L.build_layout([(c, schema[c][0], L.DT_F32, 1) for c in cols])
    plan_args = dict(num_trainers=1, num_reducers=a.reducers, batch_size=a.batch_size,
                     drop_last=False)
    for resident, passes in [("hbm", 1), ("hbm", 2), ("hbm", 4), ("hbm", 8), ("host", 1)]:
        t0 = time.perf_counter()
        opts = dict(chunk_passes=passes, resident=resident)
        if resident == "host":
            opts["stream_chunk_rows"] = a.batch_size
        eng = DeviceShuffleEngine(files, plan_args, layout_fn, 1, **opts)
        first_ms, epoch_ms, cold_s = [], [], None
        for epoch in range(a.epochs):
            buf = eng.start_epoch(epoch)[0]
            chunks = [ShuffledChunk(buf, i, lo, hi)
                      for i, (lo, hi) in enumerate(eng.plan.trainer_chunks(0))]
            chunks[0].wait()
            if epoch == 0:
                torch.cuda.current_stream().synchronize()
                cold_s = time.perf_counter() - t0
            for c in chunks[1:]:
                c.wait()
            torch.cuda.synchronize()
            if epoch >= 2:
                epoch_ms.append(eng.epoch_kernel_ms(epoch))
                first_ms.append(eng.first_pass_ms(epoch) if passes > 1 else epoch_ms[-1])
            buf.release()
        med = lambda v: sorted(v)[len(v) // 2] if v else None   # noqa: E731
        print(json.dumps({"resident": resident, "chunk_passes": eng.chunk_passes,
                          "reducer_chunks": a.reducers, "rows": a.rows_per_gpu, "cols": a.cols,
                          "first_chunk_ms": med(first_ms), "epoch_ms": med(epoch_ms),
                          "cold_first_chunk_s": cold_s, "ingest_seconds": eng.ingest_seconds}),
              flush=True)
        eng.close()


if __name__ == "__main__":
    main()
