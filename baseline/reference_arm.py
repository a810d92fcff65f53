"""``bench.py --impl reference``: the UNMODIFIED reference on the same box.

Runs ``ray_shuffling_data_loader.TorchShufflingDataset`` from ``baseline/_ref``
(pip-installed with ``--no-deps`` from ``/root/reference``; byte-identical to
upstream) through its own public API and stock code path: rank 0 constructs the
dataset (which creates the queue actor and launches ``shuffle`` as a remote
task), other ranks connect by actor name, every trainer iterates batches and -
like the reference's own example (``examples/horovod/ray_torch_shuffle.py:204-207``)
- copies each tensor to its GPU with ``.cuda()``.

None of the product's engine, kernels or dataset classes are imported here. Two
things are ours and are stated in the JSON line: the *input files* (written once
by our generator so both arms read identical Parquet) and the *Ray substrate*
(``baseline/ray_shim`` - Ray itself cannot be installed offline in this image).
If the reference cannot run, one line ``{"impl": "reference", "unavailable": ...}``
is printed and the process exits 0.
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def _unavailable(why: str):
    import bench
    bench.emit_json({"impl": "reference", "unavailable": why})
    sys.exit(0)


def main(args):
    import bench
    bench.claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    ref_dir = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "ray_shuffling_data_loader")):
        if rank == 0:
            _unavailable("baseline/_ref is missing: run `python -m pip install --no-index "
                         "--no-build-isolation --no-deps --target baseline/_ref /root/reference`")
        sys.exit(0)
    substrate = "ray"
    try:
        import ray  # noqa: F401  (a real Ray, if the image ever has one)
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "ray_shim"))
        substrate = "ray_shim (baseline/ray_shim; Ray is not installable offline)"
    sys.path.insert(0, ref_dir)
    try:
        import ray
        import torch
        import torch.distributed as dist
        from ray_shuffling_data_loader import TorchShufflingDataset
    except Exception as e:  # pragma: no cover
        if rank == 0:
            _unavailable(f"reference import failed: {type(e).__name__}: {e}")
        sys.exit(0)

    # Input files: identical bytes for both arms (generated outside timing).
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("cuda:nccl,cpu:gloo", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    bench.generate_my_share(args, rank, world)
    if world > 1:
        dist.barrier()
    _, files = bench.dataset_files(args, world)

    # Same rule as our arm (bench.effective_counts): --steps/--warmup are minimums;
    # warm up for >= max_concurrent_epochs whole epochs (those are shuffled at
    # construction), then time a whole number of epochs. The reference's shuffle
    # driver is a remote task with no handle to drain, so - unlike our arm - the
    # region does not wait for the trailing in-flight epochs (this flatters it).
    batches_per_epoch = -(-args.rows_per_gpu // args.batch_size)
    window = max(1, args.max_concurrent_epochs)
    warm_ep, timed_ep = bench.effective_counts(args.steps, args.warmup, batches_per_epoch,
                                               window, min_timed_epochs=args.min_timed_epochs)
    steps, warmup = timed_ep * batches_per_epoch, warm_ep * batches_per_epoch
    epochs = warm_ep + timed_ep + window

    if rank == 0:
        ray.init()
    if world > 1:
        dist.barrier()
    if rank != 0:
        ray.init(address="auto")

    feature_columns, label_column, _, _ = bench.schema_setup(args, torch)
    t_construct = time.perf_counter()
    if rank == 0:
        ds = TorchShufflingDataset(files, epochs, world, args.batch_size, rank,
                                   num_reducers=world, max_concurrent_epochs=window,
                                   feature_columns=feature_columns, label_column=label_column)
    if world > 1:
        dist.barrier()
    if rank != 0:
        ds = TorchShufflingDataset(files, epochs, world, args.batch_size, rank,
                                   num_reducers=world, max_concurrent_epochs=window,
                                   feature_columns=feature_columns, label_column=label_column)

    sampler = bench.ClockSampler(range(world)) if rank == 0 else None
    dev = torch.device("cuda", torch.cuda.current_device())
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d = [0]
    checksum = [0.0]
    state = {"it": None, "epoch": 0}

    def next_batch():
        while True:
            if state["it"] is None:
                ds.set_epoch(state["epoch"])
                state["it"] = iter(ds)
            try:
                return next(state["it"])
            except StopIteration:
                state["it"] = None
                state["epoch"] += 1

    def step():
        data, target = next_batch()
        # the reference example's H2D: pageable .cuda() of every tensor
        data = [t.cuda() for t in data]
        target = target.cuda()
        h2d[0] += sum(t.numel() * t.element_size() for t in data) + \
            target.numel() * target.element_size()
        # same sink as our arm: reduce every value of the batch, read it back
        nonlocal_acc = torch.stack([t.sum(dtype=torch.float64) for t in data]).sum() \
            + target.sum(dtype=torch.float64)
        acc.add_(nonlocal_acc)
        checksum[0] = float(acc.item())

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    h2d[0] = 0
    # like our arm: the clock starts BEFORE the first timed batch is fetched
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    h2d_bytes = h2d[0]
    checksum = checksum[0]
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    wall_t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
    ms, wall = float(ms.item()), float(wall_t.item())
    if rank == 0:
        rows = steps * args.batch_size * world
        value = rows / wall
        out = {
            "metric": bench.METRIC, "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": wall * 1e3 / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "float32", "data": "synthetic", "impl": "reference",
            "substrate": substrate,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "epochs_timed": timed_ep, "epochs_warmup": warm_ep,
            "batches_per_epoch": batches_per_epoch,
            "config": bench.shape_config(args, world),
            "batches_per_sec": value / args.batch_size,
            "e2e": {"value": value, "unit": "rows/s",
                    "h2d_bytes_per_step": int(h2d_bytes / steps),
                    "d2h_bytes_per_step": 8, "ms_per_step": wall * 1e3 / steps,
                    "device_ms_per_step": ms / steps},
            "gpu_launches": 0, "clocks": clocks, "checksum": checksum,
            "note": "reference has no device-resident mode: value == e2e "
                    "(host shuffle + pageable H2D every step)",
        }
        bench.emit_json(out)
    # leave quickly: the reference's shuffle driver may still be producing epochs
    if world > 1:
        dist.barrier()
    if rank == 0:
        try:
            ray.shutdown()
        except Exception:
            pass
        if not args.keep_data:
            import shutil
            d, _ = bench.dataset_files(args, world)
            shutil.rmtree(d, ignore_errors=True)
    os._exit(0)
