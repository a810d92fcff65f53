#!/usr/bin/env python
"""BASELINE.json config 5: ResNet-50 torch trainer fed by TorchShufflingDataset,
images/s end to end (the reference's Horovod example,
``examples/horovod/ray_torch_shuffle.py:143-253``, with the training step it left
commented out actually executed).

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node N \\
        benchmarks/resnet50_images.py --gpus N [--impl reference]

Both arms train the same random-init torchvision ResNet-50 (bf16 autocast,
channels-last, SGD, NCCL DDP) on the same synthetic Parquet files: 3 x 224 x 224
*uint8* images in a list column plus an int64 label, ``--images-per-gpu`` rows per
rank, re-shuffled globally every epoch.

* ours       ``TorchShufflingDataset`` with ``feature_types=[torch.uint8]``: pixels stay
             bytes across NVLink and in the epoch buffer (4x fewer bytes than float32;
             the wide scatter kernel moves 16 bytes per lane); a batch is a zero-copy
             view in HBM and the step converts it to channels-last bf16 on the fly.
* reference  the unmodified reference package (``baseline/_ref`` on ``baseline/ray_shim``):
             pandas object cells -> ``np.stack`` -> ``torch.as_tensor`` on the host, then the
             example's pageable ``.cuda()`` copies (``ray_torch_shuffle.py:204-207``).

Epoch 0 is warm-up (cold ingest, cuDNN autotune); epochs 1.. are timed with CUDA
events, max over ranks. Reported: images/s, per-epoch time, and the example's batch
wait statistics (host time between the end of a step and the arrival of the next
batch, ``:199-247``). One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--images-per-gpu", type=int, default=8192)
    p.add_argument("--image-size", type=int, default=224)
    p.add_argument("--batch-size", type=int, default=256, help="per GPU")
    p.add_argument("--epochs", type=int, default=4, help="total; the first one is warm-up")
    p.add_argument("--files-per-gpu", type=int, default=4)
    p.add_argument("--num-reducers-per-trainer", type=int, default=4)
    p.add_argument("--model", choices=["resnet50", "none"], default="resnet50",
                   help="none: loader only (the reference example's mock step without the sleep)")
    p.add_argument("--data-dir", default=os.path.join(tempfile.gettempdir(), "rsdl_resnet_bench"))
    p.add_argument("--keep-data", action="store_true")
    return p.parse_args()


def write_my_files(args, rank, world):
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    px = 3 * args.image_size * args.image_size
    d = os.path.join(args.data_dir, f"s{args.image_size}_n{args.images_per_gpu}_w{world}")
    os.makedirs(d, exist_ok=True)
    per = args.images_per_gpu // args.files_per_gpu
    files = [os.path.join(d, f"images_{r}_{i}.parquet.snappy")
             for r in range(world) for i in range(args.files_per_gpu)]
    for i in range(args.files_per_gpu):
        fn = files[rank * args.files_per_gpu + i]
        if os.path.exists(fn + ".ok"):
            continue
        rng = np.random.default_rng([77, rank, i])
        img = rng.integers(0, 256, (per, px), dtype=np.uint8)
        first = (rank * args.files_per_gpu + i) * per
        tbl = pa.table({
            "image": pa.FixedSizeListArray.from_arrays(pa.array(img.reshape(-1)), px),
            "labels": pa.array((np.arange(first, first + per) % 1000).astype(np.int64))})
        pq.write_table(tbl, fn, compression="snappy", row_group_size=max(1, per // 4))
        open(fn + ".ok", "w").close()
    return d, files


def build_step(args, torch, device, world, local_rank):
    """-> step(images_uint8_or_float[B,3,H,W], labels[B]) running fwd+bwd+SGD."""
    if args.model == "none":
        sink = torch.zeros(1, dtype=torch.float64, device=device)

        def step(img, y):
            sink.add_(img.sum(dtype=torch.float64) + y.sum())
        return step, None
    from ray_shuffling_data_loader_b200.models import build_resnet50
    import torch.nn.functional as F
    model = build_resnet50().to(device)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    mean = torch.tensor([0.485, 0.456, 0.406], device=device).view(1, 3, 1, 1) * 255
    inv_std = 1.0 / (torch.tensor([0.229, 0.224, 0.225], device=device).view(1, 3, 1, 1) * 255)

    def step(img, y):
        # bytes -> normalised bf16 channels-last, fused by the caching allocator's
        # single elementwise kernel chain; the loader delivered views, no copy before this
        x = ((img.to(torch.float32) - mean) * inv_std).to(torch.bfloat16)
        x = x.contiguous(memory_format=torch.channels_last)
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(model(x).float(), y.reshape(-1))
        loss.backward()
        opt.step()
    return step, model


def main():
    args = parse_args()
    import bench                      # stdout discipline + clock sampler
    bench.claim_stdout()
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    substrate = None
    if args.impl == "ours":
        from ray_shuffling_data_loader_b200.parallel import bootstrap
        bootstrap.init_from_env()
        from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    else:
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("cuda:nccl,cpu:gloo", rank=rank, world_size=world,
                                    device_id=device)
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "ray_shuffling_data_loader")):
            if rank == 0:
                bench.emit_json({"impl": "reference", "unavailable": "baseline/_ref is missing"})
            return
        substrate = "ray"
        try:
            import ray  # noqa: F401
        except ImportError:
            sys.path.insert(0, os.path.join(ROOT, "baseline", "ray_shim"))
            substrate = "ray_shim (baseline/ray_shim; Ray is not installable offline)"
        sys.path.insert(0, ref_dir)
        import ray
        from ray_shuffling_data_loader import TorchShufflingDataset
    d, files = write_my_files(args, rank, world)
    if world > 1:
        dist.barrier()
    s = args.image_size
    num_reducers = world * args.num_reducers_per_trainer
    kw = dict(num_reducers=num_reducers, max_concurrent_epochs=2, feature_columns=["image"],
              feature_shapes=[(3, s, s)], feature_types=[torch.uint8], label_column="labels",
              label_type=torch.int64)
    if args.impl == "ours":
        ds = TorchShufflingDataset(files, args.epochs, world, args.batch_size, rank, seed=1,
                                   backend="cuda", **kw)
    else:
        if rank == 0:
            ray.init()
        if world > 1:
            dist.barrier()
        if rank != 0:
            ray.init(address="auto")
        if rank == 0:
            ds = TorchShufflingDataset(files, args.epochs, world, args.batch_size, rank, **kw)
        if world > 1:
            dist.barrier()
        if rank != 0:
            ds = TorchShufflingDataset(files, args.epochs, world, args.batch_size, rank, **kw)
    step, model = build_step(args, torch, device, world, local_rank)
    sampler = bench.ClockSampler(range(world)) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    waits, images = [], 0
    wall0 = None
    for epoch in range(args.epochs):
        if epoch == 1:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if sampler:
                sampler.start()
            wall0 = time.perf_counter()
            ev0.record()
        ds.set_epoch(epoch)
        t_last = time.perf_counter()
        for feats, label in ds:
            if epoch >= 1:
                waits.append(time.perf_counter() - t_last)
            img = feats[0]
            if args.impl == "reference":
                img, label = img.cuda(), label.cuda()         # the example's pageable copies
            step(img, label)
            if epoch >= 1:
                images += int(label.shape[0])
            t_last = time.perf_counter()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    stats = torch.tensor([float(images), float(np.sum(waits)), float(np.max(waits))],
                         dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        mx = stats[2:].clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        stats[2] = mx[0]
    total_images = float(stats[0].item())
    timed_epochs = args.epochs - 1
    if rank == 0:
        sec = float(ms.item()) / 1e3
        nsteps = max(1, len(waits))
        out = {
            "metric": "images_per_sec", "value": total_images / sec, "unit": "images/s",
            "n_gpus": world, "impl": args.impl, "substrate": substrate,
            "epochs_timed": timed_epochs, "seconds_per_epoch": sec / timed_epochs,
            "steps_per_epoch_per_gpu": nsteps // timed_epochs,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"model": "ResNet-50 (torchvision, random init)" if args.model != "none"
                       else "loader only", "image": [3, s, s], "stored_as": "uint8",
                       "images_per_gpu": args.images_per_gpu, "batch_per_gpu": args.batch_size,
                       "global_batch": args.batch_size * world, "parallelism": f"dp{world}",
                       "num_reducers": num_reducers, "max_concurrent_epochs": 2},
            "batch_wait_mean_ms": float(stats[1].item()) / (nsteps * world) * 1e3,
            "batch_wait_max_ms": float(stats[2].item()) * 1e3,
            "batch_wait_share_of_step": float(stats[1].item()) / world / max(wall, 1e-9),
            "wall_seconds": wall, "clocks": clocks,
        }
        if args.impl == "ours":
            eng = ds.dataset.engine
            out["engine"] = {"wide_fields": len(eng.wide_field_idx), "chunk_passes": eng.chunk_passes,
                             "row_bytes": eng.layout.row_pitch,
                             "shuffle_kernel_ms": eng.epoch_kernel_ms(args.epochs - 1),
                             "first_pass_ms": eng.first_pass_ms(args.epochs - 1)}
        bench.emit_json(out)
    if args.impl == "ours":
        ds.dataset.close()
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.keep_data:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    if args.impl == "reference":
        try:
            if rank == 0:
                ray.shutdown()
        except Exception:
            pass
        os._exit(0)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
