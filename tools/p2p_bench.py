#!/usr/bin/env python
"""NVLink P2P store/copy bandwidth between GPU 0 and GPU 1 (single process):
cudaMemcpyPeer-style torch copy, and a plain vectorised store kernel via torch
(`dst.copy_(src)` with dst on the peer). Roofline reference for the scatter."""
import json
import torch


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(0)
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(0)
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    torch.cuda.set_device(0)
    n = 1 << 30
    src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    dst_local = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    dst_peer = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    out = {}
    ms = timeit(lambda: dst_local.copy_(src))
    out["local_copy_gbps_rw"] = 2 * n / ms / 1e6
    ms = timeit(lambda: dst_peer.copy_(src))
    out["peer_copy_gbps"] = n / ms / 1e6
    # random row scatter with torch index_copy_ onto the peer is not possible
    # across devices; the scatter kernel itself is measured by kernel_bench --peer.
    print(json.dumps(out))


if __name__ == "__main__":
    main()
