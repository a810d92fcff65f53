#!/usr/bin/env python
"""What can an all-to-all over NVLink 5 / NVSwitch deliver on this box, per GPU,
when EVERY GPU sends to EVERY peer at the same time? ("name the
gap".) The fused scatter kernel is judged against 900 GB/s (spec) and 774 GB/s (a
single one-directional peer copy); neither is the ceiling of a *simultaneous
bidirectional all-to-all*, which is what an epoch's shuffle is. Three references,
same bytes per peer as one epoch of the headline config (3.2 GB / N per peer):

  ce      copy engines: one cudaMemcpyAsync per peer, N-1 streams
  sm      SM stores: our row-copy kernel writing 4 KB rows round-robin to all peers
          (contiguous 512 B per warp store - the friendliest pattern for NVLink)
  nccl    torch.distributed.all_to_all_single, equal splits

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node N tools/a2a_ceiling.py

Timing: CUDA events on the device, after warm-up, barrier + synchronize on both
sides, max over ranks. Prints one JSON line per method from rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes-per-gpu", type=int, default=3_200_000_000)
    ap.add_argument("--row-bytes", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    from ray_shuffling_data_loader_b200.parallel import bootstrap
    from ray_shuffling_data_loader_b200 import _C as C
    import torch.distributed._symmetric_memory as symm_mem
    ctx = bootstrap.init_from_env()
    rank, world = ctx.rank, ctx.world
    dev = torch.device("cuda", ctx.local_rank)
    torch.cuda.set_device(dev)
    C.set_device(ctx.local_rank)
    pitch = a.row_bytes
    rows_per_peer = a.bytes_per_gpu // world // pitch
    blk = rows_per_peer * pitch                       # bytes to each destination (incl. self)
    # destination: symmetric allocation [world][blk]: block s receives from source s
    t = symm_mem.empty((world * blk,), dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.zero_()
    peers = [int(p) for p in hdl.buffer_ptrs]
    src = torch.empty((world * blk,), dtype=torch.uint8, device=dev)
    src.view(torch.int32).random_(0, 1 << 30)
    torch.cuda.synchronize()
    hdl.barrier()

    def timed(fn, sync_all_streams=None):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(ms.item())

    def report(name, ms, note=""):
        if rank == 0:
            egress = (world - 1) * blk / (ms / 1e3) / 1e9
            print(json.dumps({"method": name, "n_gpus": world, "bytes_per_peer": blk,
                              "ms": ms, "egress_gbps_per_gpu": egress,
                              "frac_of_900": egress / 900.0, "note": note}), flush=True)

    # ---- copy engines: one async copy per peer, each on its own stream, joined on the main one
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]

    def ce():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for p in range(world):
            if p == rank:
                continue
            s = streams[p]
            s.wait_event(ev)
            C.memcpy_async(peers[p] + rank * blk, src.data_ptr() + p * blk, blk, C.D2D,
                           s.cuda_stream)
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)
    report("ce", timed(ce), "cudaMemcpyAsync per peer on its own stream")

    # ---- SM stores: one launch, rows dealt round-robin to the peers
    base = min(peers)
    n_rows = rows_per_peer * (world - 1)
    others = [p for p in range(world) if p != rank]
    i = torch.arange(n_rows, dtype=torch.int64, device=dev)
    which = i % (world - 1)
    k = i // (world - 1)
    peer_base = torch.tensor([peers[p] - base for p in others], dtype=torch.int64, device=dev)
    src_blk = torch.tensor([p * blk for p in others], dtype=torch.int64, device=dev)
    dst_off = peer_base[which] + rank * blk + k * pitch
    src_idx = (src_blk[which] + k * pitch) // pitch
    stream = torch.cuda.current_stream().cuda_stream

    def sm():
        C.place_rows(src.data_ptr(), dst_off.data_ptr(), n_rows, pitch, base, stream,
                     src_idx=src_idx.data_ptr())
    report("sm", timed(sm), f"place_rows kernel, {pitch}-byte rows round-robin over peers")

    # ---- NCCL
    recv = torch.empty_like(src)

    def nccl():
        dist.all_to_all_single(recv, src)
    report("nccl", timed(nccl), "all_to_all_single equal splits (includes the local block)")
    hdl.barrier()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
