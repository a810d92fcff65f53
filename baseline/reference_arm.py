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
    t_process = time.perf_counter()
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
                                   num_reducers=world * args.reducers_per_trainer,
                                   max_concurrent_epochs=window,
                                   feature_columns=feature_columns, label_column=label_column)
    if world > 1:
        dist.barrier()
    if rank != 0:
        ds = TorchShufflingDataset(files, epochs, world, args.batch_size, rank,
                                   num_reducers=world * args.reducers_per_trainer,
                                   max_concurrent_epochs=window,
                                   feature_columns=feature_columns, label_column=label_column)

    sampler = bench.ClockSampler(range(world)) if rank == 0 else None
    dev = torch.device("cuda", torch.cuda.current_device())
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counters = {"h2d": 0, "rows": 0, "steps": 0}
    checksum = [0.0]

    def consume_epoch(epoch):
        """One whole epoch through the reference's public iterator, consumed the way
        its own example does (examples/horovod/ray_torch_shuffle.py:202-207)."""
        ds.set_epoch(epoch)
        for data, target in ds:
            # the reference example's H2D: pageable .cuda() of every tensor
            data = [t.cuda() for t in data]
            target = target.cuda()
            counters["h2d"] += sum(t.numel() * t.element_size() for t in data) + \
                target.numel() * target.element_size()
            counters["rows"] += int(target.shape[0])
            counters["steps"] += 1
            # same sink as our arm: reduce every value of the batch, read it back
            acc.add_(torch.stack([t.sum(dtype=torch.float64) for t in data]).sum()
                     + target.sum(dtype=torch.float64))
            checksum[0] = float(acc.item())

    def agree(x):
        """max over ranks of a host float (so every rank takes the same decision)"""
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up: whole epochs, each one timed - the last one is the fallback measurement
    # if the time budget (the driver kills a run after ~870 s) leaves no room for more
    last_warm = None
    for e in range(warm_ep):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c0 = dict(counters)
        t0 = time.perf_counter()
        consume_epoch(e)
        torch.cuda.synchronize()
        last_warm = {"wall": agree(time.perf_counter() - t0), "rows": counters["rows"] - c0["rows"],
                     "steps": counters["steps"] - c0["steps"], "h2d": counters["h2d"] - c0["h2d"]}
    elapsed = agree(time.perf_counter() - t_process)
    room = args.time_budget_s - elapsed
    fit = int(room // max(last_warm["wall"], 1e-3))
    fallback = fit < 1
    if not fallback:
        timed_ep = max(1, min(timed_ep, fit))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    if fallback:
        measured = dict(last_warm)
        measured["ms"] = last_warm["wall"] * 1e3
        timed_ep = 1
    else:
        c0 = dict(counters)
        wall0 = time.perf_counter()
        ev0.record()
        for e in range(warm_ep, warm_ep + timed_ep):
            consume_epoch(e)
        ev1.record()
        torch.cuda.synchronize()
        measured = {"wall": agree(time.perf_counter() - wall0), "ms": agree(ev0.elapsed_time(ev1)),
                    "rows": counters["rows"] - c0["rows"], "steps": counters["steps"] - c0["steps"],
                    "h2d": counters["h2d"] - c0["h2d"]}
    clocks = sampler.stop() if sampler else None
    tot = torch.tensor([float(measured["rows"]), float(measured["steps"]), float(measured["h2d"])],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    rows_all, steps_all, h2d_all = (float(x) for x in tot.tolist())
    wall, ms = measured["wall"], measured["ms"]
    checksum = checksum[0]
    if rank == 0:
        steps = max(1, int(round(steps_all / world)))        # batches per trainer in the region
        warmup = warm_ep * batches_per_epoch
        value = rows_all / wall
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
            "timed_region": ("last warm-up epoch (time budget of %.0f s left no room)"
                             % args.time_budget_s) if fallback else "whole epochs after warm-up",
            "e2e": {"value": value, "unit": "rows/s",
                    "h2d_bytes_per_step": int(h2d_all / max(1.0, steps_all)),
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
