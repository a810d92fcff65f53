#!/usr/bin/env python
"""Pinned host -> device copy bandwidth: the roofline of the end-to-end
(streaming) path. One flat 64 MB copy vs the engine's 2-D chunk copy
(64 columns x 1 MB, host stride = whole column)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from ray_shuffling_data_loader_b200 import _C
from ray_shuffling_data_loader_b200.runtime.device_engine import pinned_array


def main():
    torch.cuda.set_device(0)
    _C.set_device(0)
    cols, rows_total, chunk = 64, 12_500_000, 250_000
    stride = (rows_total * 4 + 255) // 256 * 256
    host, hptr = pinned_array(_C, (cols * stride,), np.uint8)
    host[:] = 1
    dev = _C.device_malloc(cols * chunk * 4 + 4096)
    stream = _C.stream_create(0)
    out = {}
    for name in ("flat_64MB", "2d_64x1MB"):
        times = []
        for it in range(12):
            e0, e1 = _C.event_create(True), _C.event_create(True)
            _C.event_record(e0, stream)
            for rep in range(4):
                if name == "flat_64MB":
                    _C.memcpy_async(dev, hptr, cols * chunk * 4, _C.H2D, stream)
                else:
                    _C.memcpy2d_async(dev, chunk * 4, hptr + rep * chunk * 4, stride, chunk * 4,
                                      cols, _C.H2D, stream)
            _C.event_record(e1, stream)
            _C.stream_synchronize(stream)
            times.append(_C.event_elapsed_ms(e0, e1) / 4)
        best = min(times[2:])
        out[name] = {"ms": best, "gbps": cols * chunk * 4 / best / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
