#!/usr/bin/env python
"""Tail fields A/B on one GPU: a DLRM-style row (the reference's DATA_SPEC: key + 19 int64
features kept as int64, float64 label delivered as float32) shuffled with the label folded
into the fast kernel (``tail_fields=True``, one launch per epoch) vs written by the generic
kernel (``tail_fields=False``, two launches). Prints one JSON line per setting."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import argparse
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--batch-size", type=int, default=250_000)
    ap.add_argument("--schema", default="dataspec")
    ap.add_argument("--data-dir", default=os.environ.get("RSDL_BENCH_DIR", "/tmp/rsdl_bench"))
    a = ap.parse_args()
    import bench
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    torch.cuda.set_device(0)
    bench.generate_my_share(a, 0, 1)
    _, files = bench.dataset_files(a, 1)

    def layout_fn(schema):
        feats = [c for c in schema if c != "labels"]
        return L.build_layout([(c, schema[c][0], schema[c][0], 1) for c in feats]
                              + [("labels", schema["labels"][0], L.DT_F32, 1)])
    plan_args = dict(num_trainers=1, num_reducers=1, batch_size=a.batch_size, drop_last=False)
    for tails in (True, False, True, False):
        eng = DeviceShuffleEngine(files, plan_args, layout_fn, 1, tail_fields=tails)
        ms = []
        for epoch in range(8):
            buf = eng.start_epoch(epoch)[0]
            buf.wait()
            torch.cuda.synchronize()
            if epoch >= 2:
                ms.append(eng.epoch_kernel_ms(epoch))
            buf.release()
        ms = sorted(m for m in ms if m)
        print(json.dumps({"tail_fields": tails, "row_pitch": eng.layout.row_pitch,
                          "fast_mode": eng.fast_mode, "tail_fields_folded": len(eng.tail_field_idx),
                          "generic_runs": len(eng.generic_runs),
                          "launches_per_epoch": 1 + len(eng.generic_runs),
                          "epoch_ms_median": ms[len(ms) // 2], "epoch_ms_min": ms[0]}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
