#!/usr/bin/env python
"""Distributed training fed by ``TorchShufflingDataset`` (component C17).

The B200 counterpart of the reference's Horovod example
(``examples/horovod/ray_torch_shuffle.py``): same data-generation and loader
flags, same per-epoch / overall **batch wait time** statistics
(``:199-247``), and ``--mock-train-step-time`` to reproduce its sleep-only
"training". Differences: workers are torchrun ranks with NCCL DDP instead of
Horovod-on-Ray; batches arrive as CUDA tensors (no ``.cuda()`` copies, reference
``:204-207``); and unless a mock step time is given a real model is trained:

    --model mlp       TabularMLP on N float32 columns (default)
    --model dlrm      embedding net on the reference's DATA_SPEC schema
    --model resnet50  images stored as a list column, ResNet-50 (images/s)

    torchrun --standalone --nproc-per-node 8 examples/ddp/torch_shuffle.py --epochs 3
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import tempfile
import time
import timeit

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from ray_shuffling_data_loader_b200 import TorchShufflingDataset  # noqa: E402
from ray_shuffling_data_loader_b200.data_generation import (DATA_SPEC, float_spec,  # noqa: E402
                                                            generate_data)
from ray_shuffling_data_loader_b200.parallel import bootstrap  # noqa: E402
from ray_shuffling_data_loader_b200.stats import human_readable_size  # noqa: E402

DEFAULT_DATA_DIR = os.path.join(tempfile.gettempdir(), "rsdl_example_data")

parser = argparse.ArgumentParser(description="Shuffling data loader + DDP example")
parser.add_argument("--batch-size", type=int, default=250000)
parser.add_argument("--epochs", type=int, default=10)
parser.add_argument("--lr", type=float, default=0.01)
parser.add_argument("--momentum", type=float, default=0.5)
parser.add_argument("--no-cuda", action="store_true", default=False)
parser.add_argument("--seed", type=int, default=42)
parser.add_argument("--log-interval", type=int, default=10)
parser.add_argument("--mock-train-step-time", type=float, default=None,
                    help="sleep this long instead of training (the reference always did)")
parser.add_argument("--model", choices=["mlp", "dlrm", "resnet50"], default="mlp")
parser.add_argument("--bf16", action="store_true", help="deliver features as bf16")
# Synthetic training data generation settings.
parser.add_argument("--cache-files", action="store_true", default=False)
parser.add_argument("--num-rows", type=int, default=2 * (10**7))
parser.add_argument("--num-files", type=int, default=25)
parser.add_argument("--num-columns", type=int, default=64)
parser.add_argument("--max-row-group-skew", type=float, default=0.0)
parser.add_argument("--num-row-groups-per-file", type=int, default=5)
parser.add_argument("--data-dir", type=str, default=DEFAULT_DATA_DIR)
parser.add_argument("--image-size", type=int, default=64)
# Shuffling data loader settings.
parser.add_argument("--num-reducers", type=int, default=32)
parser.add_argument("--max-concurrent-epochs", type=int, default=2)


def make_image_files(args):
    """ResNet-50 config: images as a fixed-size list column + int64 label."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    os.makedirs(args.data_dir, exist_ok=True)
    rng = np.random.default_rng(args.seed)
    per = args.num_rows // args.num_files
    px = 3 * args.image_size * args.image_size
    files = []
    for i in range(args.num_files):
        img = rng.integers(0, 256, (per, px), dtype=np.uint8)     # pixels are stored as bytes
        tbl = pa.table({
            "image": pa.FixedSizeListArray.from_arrays(pa.array(img.reshape(-1)), px),
            "labels": pa.array(rng.integers(0, 1000, per, dtype=np.int64))})
        fn = os.path.join(args.data_dir, f"images_{i}.parquet.snappy")
        pq.write_table(tbl, fn, compression="snappy", row_group_size=max(1, per // 4))
        files.append(fn)
    return files, args.num_files * per * (px + 8)


def get_files(args, rank):
    cache_path = os.path.join(tempfile.gettempdir(), f"data_cache_{args.model}")
    if args.cache_files and os.path.exists(cache_path):
        try:
            with open(cache_path, "rb") as f:
                return pickle.load(f)
        except Exception as exc:
            print(f"Cache load failed - {exc}")
    result = None
    if rank == 0:
        print(f"Generating {args.num_rows} rows over {args.num_files} files, with "
              f"{args.num_row_groups_per_file} row groups per file.")
        if args.model == "resnet50":
            result = make_image_files(args)
        else:
            spec = DATA_SPEC if args.model == "dlrm" else float_spec(args.num_columns, np.float32)
            result = generate_data(args.num_rows, args.num_files, args.num_row_groups_per_file,
                                   args.max_row_group_skew, args.data_dir, data_spec=spec,
                                   seed=args.seed)
        if args.cache_files:
            with open(cache_path, "wb") as f:
                pickle.dump(result, f)
    return bootstrap.broadcast_object(result, 0)


def create_dataset(args, filenames, rank, world_size):
    print(f"Creating Torch shuffling dataset for worker {rank} with "
          f"{args.batch_size} batch size, {args.epochs} epochs, {args.num_reducers} "
          f"reducers, and {world_size} trainers.")
    common = dict(num_reducers=args.num_reducers, max_concurrent_epochs=args.max_concurrent_epochs,
                  seed=args.seed, backend="cpu" if args.no_cuda else None)
    if args.model == "dlrm":
        cols = list(DATA_SPEC.keys())
        label = cols.pop()
        return TorchShufflingDataset(filenames, args.epochs, world_size, args.batch_size, rank,
                                     feature_columns=cols, feature_types=[torch.int64] * len(cols),
                                     label_column=label, label_type=torch.float32, **common)
    if args.model == "resnet50":
        s = args.image_size
        return TorchShufflingDataset(filenames, args.epochs, world_size, args.batch_size, rank,
                                     feature_columns=["image"], feature_shapes=[(3, s, s)],
                                     # bytes stay bytes through the shuffle (4x fewer NVLink
                                     # bytes than float32); --bf16 converts inside the kernel
                                     feature_types=[torch.bfloat16 if args.bf16 else torch.uint8],
                                     label_column="labels", label_type=torch.int64, **common)
    cols = [f"f{i}" for i in range(args.num_columns - 1)]
    dt = torch.bfloat16 if args.bf16 else torch.float32
    return TorchShufflingDataset(filenames, args.epochs, world_size, args.batch_size, rank,
                                 feature_columns=cols, feature_types=[dt] * len(cols),
                                 label_column="labels", label_type=torch.float32,
                                 packed_features=True, **common)


def build_model(args, device):
    from ray_shuffling_data_loader_b200 import models
    if args.model == "dlrm":
        card = {c: hi for c, (lo, hi, dt) in DATA_SPEC.items() if c != "labels"}
        return models.EmbeddingTabularNet(card).to(device)
    if args.model == "resnet50":
        return models.build_resnet50().to(device)
    return models.TabularMLP(args.num_columns - 1).to(device)


def train_main(args):
    ctx = bootstrap.init_from_env()
    rank, world = ctx.rank, ctx.world
    torch.manual_seed(args.seed)
    use_cuda = torch.cuda.is_available() and not args.no_cuda
    device = torch.device("cuda", ctx.local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    filenames, num_bytes = get_files(args, rank)
    if rank == 0:
        print(f"Generated {len(filenames)} files totalling {human_readable_size(num_bytes)}.")
    train_dataset = create_dataset(args, filenames, rank, world)
    mock = args.mock_train_step_time
    model = optimizer = None
    if mock is None:
        model = build_model(args, device)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(
                model, device_ids=[ctx.local_rank] if use_cuda else None)
            if use_cuda and os.environ.get("RSDL_EXAMPLE_GRAD_COMPRESSION") == "bf16":
                # the reference's --fp16-allreduce (hvd.Compression.fp16, :188-190)
                from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                model.register_comm_hook(None, default_hooks.bf16_compress_hook)
        optimizer = torch.optim.SGD(model.parameters(), lr=args.lr * world,
                                    momentum=args.momentum)

    def _train(epoch):
        if model is not None:
            model.train()
        train_dataset.set_epoch(epoch)
        start_epoch = timeit.default_timer()
        last_batch_time = start_epoch
        batch_wait_times, samples = [], 0
        for batch_idx, (data, target) in enumerate(train_dataset):
            batch_wait_times.append(timeit.default_timer() - last_batch_time)
            samples += target.shape[0]
            if batch_idx % args.log_interval == 0:
                print(f"Processing batch {batch_idx} in epoch {epoch} on worker {rank}.")
            if mock is not None:
                time.sleep(mock)
            else:
                optimizer.zero_grad(set_to_none=True)
                with torch.autocast(device.type, dtype=torch.bfloat16, enabled=use_cuda):
                    if args.model == "resnet50":
                        x = data[0].to(torch.bfloat16).mul_(1.0 / 255).contiguous(
                            memory_format=torch.channels_last)
                        out = model(x)
                        loss = F.cross_entropy(out.float(), target.reshape(-1))
                    elif args.model == "dlrm":
                        loss = F.binary_cross_entropy_with_logits(model(data).float(), target)
                    else:
                        loss = F.mse_loss(model(data).float(), target)
                loss.backward()
                optimizer.step()
            last_batch_time = timeit.default_timer()
        if use_cuda:
            torch.cuda.synchronize()
        epoch_duration = timeit.default_timer() - start_epoch
        w = np.asarray(batch_wait_times)
        print(f"\nEpoch {epoch}, worker {rank} stats over {len(w)} steps: "
              f"{epoch_duration:.3f}s, {samples / epoch_duration:.1f} samples/s")
        print(f"Mean batch wait time: {w.mean():.6f}s +- {w.std()}")
        print(f"Max batch wait time: {w.max():.6f}s")
        print(f"Min batch wait time: {w.min():.6f}s")
        return batch_wait_times

    print(f"Starting training on worker {rank}.")
    batch_wait_times = []
    for epoch in range(args.epochs):
        batch_wait_times.extend(_train(epoch))
    batch_wait_times.pop(0)
    print(f"Done training on worker {rank}.")
    w = np.asarray(batch_wait_times)
    print(f"\nWorker {rank} training stats over {args.epochs} epochs:")
    print(f"Mean batch wait time: {w.mean():.6f}s +- {w.std()}")
    print(f"Max batch wait time: {w.max():.6f}s")
    print(f"Min batch wait time: {w.min():.6f}s")
    # No "rank 0 must outlive the others" sleep (reference :248-253): ranks are
    # symmetric and the dataset tears down behind a barrier.
    train_dataset.dataset.close()
    bootstrap.barrier()


if __name__ == "__main__":
    train_main(parser.parse_args())
    print("Done consuming batches.")
