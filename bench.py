#!/usr/bin/env python
"""Headline benchmark: shuffled-batch delivery throughput (rows/s, batches/s).

Metric and config are BASELINE.json's: ``TorchShufflingDataset``, N trainers x N
reducers, 64 float32 columns, batch_size 250 000, weak scaling with 1.25e7 rows
(3.2 GB) per GPU (= the "1e8 rows x 64 f32 on 8xB200" table at N=8), synthetic
snappy Parquet written by ``data_generation.py``.

    python bench.py --gpus N --steps K --warmup W            # ours
    python bench.py --impl reference --gpus N --steps K ...  # unmodified reference

A *step* is one 250 000-row batch delivered to every trainer and fully consumed
(a device reduction over every byte of the batch). Two numbers are reported:

* ``value``   device-timed (CUDA events, max over ranks) rows/s with the decoded
              table resident in HBM: per-epoch fused scatter kernel + consumer.
* ``e2e``     the same metric through the public API with the table in *pinned
              host memory*: every step pays the H2D copy of one batch worth of
              source rows (64 MB) and a D2H read of the step's result (8 B).

``--steps`` / ``--warmup`` are minimums (see ``effective_counts``): the unit of
work of a shuffling loader is an epoch, and the first ``max_concurrent_epochs``
epochs are shuffled at construction, so both arms warm up for at least that many
whole epochs and then time a whole number of epochs. In our arm the timed region
opens and closes with a drained pipeline: exactly as many epochs are *shuffled*
inside it (scatter kernels, and H2D copies in the e2e phase - both counted and
reported) as are consumed. Every timed epoch is also checked for exactly-once
delivery on all ranks (all-reduced fp64 sum of every delivered value == the same
sum over the source table computed with plain torch).

Timing rules followed: W >= 3 warm-up steps, inputs (3.2 GB/epoch/GPU) far
larger than L2, CUDA events on the consumer stream bracketed by barrier +
synchronize, max over ranks, nvidia-smi clock samples during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ROWS_PER_GPU = 12_500_000
NUM_COLS = 64
BATCH_SIZE = 250_000
FILES_PER_GPU = 5
ROW_GROUPS_PER_FILE = 5
METRIC = "rows_per_sec"


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: move the real stdout aside and
    point fd 1 at stderr so that library chatter (e.g. NCCL's version banner,
    printed with printf) cannot interleave with it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr
    return _JSON_OUT


def emit_json(obj):
    out = claim_stdout()
    out.write(json.dumps(obj) + "\n")
    out.flush()


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    # Defaults: warm up for 2 whole epochs (50 batches each) so the epoch ring is
    # in steady state - nothing consumed in the timed region was shuffled before
    # it started - then time 4 whole epochs.
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=100)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    p.add_argument("--cols", type=int, default=NUM_COLS)
    p.add_argument("--batch-size", type=int, default=BATCH_SIZE)
    p.add_argument("--data-dir",
                   default=os.environ.get("RSDL_BENCH_DIR",
                                          os.path.join(tempfile.gettempdir(), "rsdl_bench")))
    p.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p")
    p.add_argument("--feature-dtype", choices=["float32", "bfloat16", "fp8"], default="float32")
    p.add_argument("--peer-alloc", choices=["symm", "ipc"], default=None)
    p.add_argument("--wait-mode", choices=["stream", "host"], default=None,
                   help="how a trainer waits for an epoch's produced flags: a wait kernel on its "
                        "own stream (no host round trip; the engine default) or a host poll")
    p.add_argument("--schema", choices=["f32", "dataspec"], default="f32",
                   help="f32: --cols float32 columns (BASELINE config 2/4); dataspec: the "
                        "reference's own DATA_SPEC table (key + 19 int64 + 1 float64, 168 B/row)")
    p.add_argument("--row-align", type=int, default=None,
                   help="pad the packed row pitch to a multiple of this (e.g. 128); 0 = never; "
                        "default = the dataset's auto rule (96..127-byte rows -> 128)")
    p.add_argument("--min-timed-epochs", type=int, default=None,
                   help="time at least this many whole epochs (default: ours 20, reference 1)")
    p.add_argument("--max-concurrent-epochs", type=int, default=2,
                   help="epoch window (BASELINE config 3 compares 1 vs 2)")
    p.add_argument("--reducers-per-trainer", type=int, default=1,
                   help="reducer chunks per trainer (BASELINE config: 1). With more, the "
                        "engine delivers an epoch in destination-chunk passes (K7) and "
                        "`first_chunk_ms` reports when the first chunk became consumable")
    p.add_argument("--chunk-passes", type=int, default=None)
    p.add_argument("--shuffle-priority", choices=["low", "high"], default=None)
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--time-budget-s", type=float, default=600.0,
                   help="reference arm: shrink the timed region (whole epochs, >= 1) so the "
                        "run ends within this many seconds of wall clock")
    p.add_argument("--keep-data", action="store_true")
    return p.parse_args()


# ---------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------

def dataset_files(args, world):
    n_files = FILES_PER_GPU * world
    tag = f"c{args.cols}" if args.schema == "f32" else args.schema
    d = os.path.join(args.data_dir, f"r{args.rows_per_gpu}_{tag}_w{world}")
    return d, [os.path.join(d, f"input_data_{i}.parquet.snappy") for i in range(n_files)]


def generate_my_share(args, rank, world):
    """Every rank writes FILES_PER_GPU files of the global table (outside any
    timed region). Global row index = position in the concatenation."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from ray_shuffling_data_loader_b200.data_generation import float_spec, generate_file
    d, files = dataset_files(args, world)
    os.makedirs(d, exist_ok=True)
    spec = float_spec(args.cols, np.float32) if args.schema == "f32" else None   # None: DATA_SPEC
    rows_per_file = args.rows_per_gpu // FILES_PER_GPU
    mine = range(rank * FILES_PER_GPU, (rank + 1) * FILES_PER_GPU)
    todo = [i for i in mine if not os.path.exists(files[i] + ".ok")]

    def one(i):
        rows = rows_per_file + (args.rows_per_gpu - rows_per_file * FILES_PER_GPU
                                if i % FILES_PER_GPU == FILES_PER_GPU - 1 else 0)
        start = (i // FILES_PER_GPU) * args.rows_per_gpu + (i % FILES_PER_GPU) * rows_per_file
        generate_file(i, start, rows, ROW_GROUPS_PER_FILE, d, spec,
                      np.random.SeedSequence([1234, i]), include_key=args.schema != "f32")
        open(files[i] + ".ok", "w").close()
    if todo:
        with ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(one, todo))
    return files


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpus = set(gpu_indices)
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for p in self.samples:
            try:
                if int(p[0]) not in self.gpus:
                    continue
                sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax),
                "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# step / epoch accounting shared by both arms
# ---------------------------------------------------------------------------

def effective_counts(steps, warmup, batches_per_epoch, window, min_timed_epochs=1):
    """``--steps`` / ``--warmup`` are *minimums*. The loader's unit of work is an
    epoch (one shuffle of the table feeds ``batches_per_epoch`` steps) and the first
    ``window`` epochs are shuffled at construction, so a clock around fewer steps
    than that would time a consumer loop on pre-shuffled data. Both arms therefore
    warm up for >= ``window`` whole epochs and time a whole number of epochs.
    -> (warm_epochs, timed_epochs)"""
    warm = max(window, -(-max(0, warmup) // batches_per_epoch))
    timed = max(min_timed_epochs, -(-max(1, steps) // batches_per_epoch))
    return warm, timed


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def run_phase(ds, engine, torch, dist, world, warm_epochs, timed_epochs, d2h_each_step):
    """Consume ``warm_epochs`` then ``timed_epochs`` whole epochs through the
    public iterator.

    Timed region (barrier + synchronize on both sides): it opens with a drained
    pipeline - the ``window`` epochs in flight are completely shuffled - and it
    closes only after the shuffles of the ``window`` epochs *following* the last
    consumed one have landed on this rank. So exactly ``timed_epochs`` epochs are
    consumed AND exactly ``timed_epochs`` epochs are shuffled (scatter kernels,
    H2D copies in host mode) inside it; nothing is consumed for free.

    Returns a dict: device ms (max over ranks), wall s, launch / byte counters of
    the region and the per-epoch fp64 sums of everything that was consumed."""
    row_pitch = engine.layout.row_pitch
    window = engine.window
    dev = torch.device("cuda", torch.cuda.current_device())
    n_ep = warm_epochs + timed_epochs
    sums = torch.zeros(n_ep, dtype=torch.float64, device=dev)
    host_acc = torch.zeros(1, dtype=torch.float64).pin_memory()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    checksum = [0.0]
    steps_done = [0]

    def consume_epoch(epoch):
        acc = sums[epoch:epoch + 1]
        ds.set_epoch(epoch)
        for features, label in ds:                # the public API: (features, label)
            # consume: every byte of the batch (features + label share one packed
            # row; the feature view starts at the batch's first byte) is read by
            # our sink kernel and reduced in fp64
            base = features[0] if isinstance(features, tuple) else features
            engine.batch_sum_all(base, acc, nbytes=base.shape[0] * row_pitch)
            if d2h_each_step:
                host_acc.copy_(acc, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                checksum[0] = float(host_acc[0])
            steps_done[0] += 1

    for e in range(warm_epochs):
        consume_epoch(e)
    warm_steps = steps_done[0]
    if not engine.wait_epochs_started(warm_epochs + window, 120.0):
        raise RuntimeError("shuffle driver did not start the in-flight epochs")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    c0 = (engine.launches, engine.scatter_launches, engine.h2d_bytes_enqueued)
    wall0 = time.perf_counter()
    ev0.record()
    for e in range(warm_epochs, n_ep):
        consume_epoch(e)
    host_loop = time.perf_counter() - wall0       # host time to ENQUEUE the region's work
    # close the books: the next ``window`` epochs must be shuffled inside the region
    if not engine.wait_epochs_started(n_ep + window, 120.0):
        raise RuntimeError("shuffle driver did not start the trailing epochs")
    for e in range(n_ep, n_ep + window):
        engine.enqueue_wait_produced(e)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    c1 = (engine.launches, engine.scatter_launches, engine.h2d_bytes_enqueued)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    engine.check_error()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    wall_t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    kernel_ms = [engine.epoch_kernel_ms(e) for e in range(warm_epochs + window, n_ep + window)]
    kernel_ms = sorted(k for k in kernel_ms if k)
    first_ms = sorted(m for m in (engine.first_pass_ms(e)
                                  for e in range(warm_epochs + window, n_ep + window)) if m)
    # the same statistic on every rank: an epoch is only as fast as the slowest source
    med = torch.tensor([kernel_ms[len(kernel_ms) // 2] if kernel_ms else 0.0],
                       dtype=torch.float64, device=dev)
    per_rank = [med.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, med)
    kernel_ms_per_rank = [float(x.item()) for x in per_rank]
    return {"ms": float(ms.item()), "wall": float(wall_t.item()),
            "steps": steps_done[0] - warm_steps, "warm_steps": warm_steps,
            "host_loop_s": host_loop,
            "launches": c1[0] - c0[0], "scatter_launches": c1[1] - c0[1],
            "h2d_bytes": c1[2] - c0[2], "epoch_sums": sums.cpu().tolist(),
            "kernel_ms": kernel_ms, "kernel_ms_per_rank": kernel_ms_per_rank,
            "first_pass_ms": first_ms[len(first_ms) // 2] if first_ms else None,
            "checksum": checksum[0] if d2h_each_step else None}


def expected_table_sum(engine, torch, dist, world):
    """Ground truth for the exactly-once check, computed with plain torch from the
    HBM-resident *source* columns (never touched by our kernels): the fp64 sum of
    every value as the loader is asked to deliver it (fp32)."""
    total = torch.zeros(1, dtype=torch.float64, device=torch.device("cuda", engine.device_index))
    for f, col in engine.source_column_tensors():
        total += col.to(torch.float32).sum(dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return float(total.item())


def exactly_once(epoch_sums, expected, rel_tol=1e-9):
    """Every epoch's all-reduced sum over every delivered value must equal the
    table's. (A lost, duplicated or torn row changes the sum; the random fp32
    payload makes an accidental match impossible in practice.)"""
    worst = max((abs(s - expected) / max(1.0, abs(expected)) for s in epoch_sums), default=0.0)
    return {"ok": bool(worst <= rel_tol), "epochs_checked": len(epoch_sums),
            "expected_sum": expected, "max_rel_err": worst, "rel_tol": rel_tol}


def schema_setup(args, torch):
    """(feature columns, label column, feature dtype, source bytes per row)."""
    dt = {"float32": torch.float32, "bfloat16": torch.bfloat16,
          "fp8": getattr(torch, "float8_e4m3fn", None)}[args.feature_dtype]
    if args.schema == "dataspec":
        from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
        names = ["key"] + list(DATA_SPEC.keys())
        return names[:-1], names[-1], dt, 8 * len(names)
    return [f"f{i}" for i in range(args.cols - 1)], "labels", dt, 4 * args.cols


def make_dataset(args, files, rank, world, epochs, resident, torch, seed=20260921):
    """The flagship public API, exactly as a user would call it."""
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    feature_columns, label_column, dt, _ = schema_setup(args, torch)
    opts = dict(resident=resident, exchange=args.exchange)
    if args.wait_mode:
        opts["wait_mode"] = args.wait_mode
    if args.peer_alloc:
        opts["peer_alloc"] = args.peer_alloc
    if args.row_align is not None:
        opts["row_align"] = args.row_align
    if resident == "host":
        opts["stream_chunk_rows"] = args.batch_size
    if args.chunk_passes is not None:
        opts["chunk_passes"] = args.chunk_passes
    if args.shuffle_priority:
        opts["shuffle_priority"] = args.shuffle_priority
    fp8 = args.feature_dtype == "fp8"
    return TorchShufflingDataset(
        files, epochs, world, args.batch_size, rank,
        num_reducers=world * args.reducers_per_trainer,
        max_concurrent_epochs=args.max_concurrent_epochs, feature_columns=feature_columns,
        feature_types=[dt] * len(feature_columns), label_column=label_column,
        label_type=dt if not fp8 else torch.float32, packed_features=True,
        fp8_block_scale=fp8, seed=seed, backend="cuda", queue_name=f"bench-{resident}", **opts)


def shape_config(args, world):
    """The benchmark *shape*: identical keys and values in both arms."""
    return {"model": f"TorchShufflingDataset {world} trainers x "
                     f"{world * args.reducers_per_trainer} reducers",
            "global_batch": args.batch_size * world,
            "rows": args.rows_per_gpu * world, "cols": args.cols if args.schema == "f32" else 21,
            "schema": args.schema, "batch_size": args.batch_size, "seq_len": None,
            "parallelism": f"dp{world}", "max_concurrent_epochs": args.max_concurrent_epochs,
            "l2_policy": "inputs larger than L2 (every epoch streams the whole per-GPU table)"}


def run_ours(args):
    claim_stdout()
    import torch
    import torch.distributed as dist
    from ray_shuffling_data_loader_b200.parallel import bootstrap
    ctx = bootstrap.init_from_env()
    rank, world = ctx.rank, ctx.world
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun")
    torch.cuda.set_device(ctx.local_rank if world > 1 else 0)
    t_gen = time.perf_counter()
    generate_my_share(args, rank, world)
    if world > 1:
        dist.barrier()
    gen_s = time.perf_counter() - t_gen
    _, files = dataset_files(args, world)
    batches_per_epoch = -(-args.rows_per_gpu // args.batch_size)
    window = max(1, args.max_concurrent_epochs)
    warm_ep, timed_ep = effective_counts(args.steps, args.warmup, batches_per_epoch, window,
                                         min_timed_epochs=args.min_timed_epochs)
    epochs = warm_ep + timed_ep + window
    sampler = ClockSampler(range(world)) if rank == 0 else None
    _, _, _, src_row_bytes = schema_setup(args, torch)

    # ---- device-timed, HBM-resident -----------------------------------------
    t_first = time.perf_counter()
    ds = make_dataset(args, files, rank, world, epochs, "hbm", torch)
    engine = ds.dataset.engine
    if sampler:
        sampler.start()
    res = run_phase(ds, engine, torch, dist, world, warm_ep, timed_ep, d2h_each_step=False)
    clocks = sampler.stop() if sampler else None
    ingest_s = getattr(engine, "ingest_seconds", None)
    check_sum = args.feature_dtype == "float32"
    expected = expected_table_sum(engine, torch, dist, world) if check_sum else None
    once = exactly_once(res["epoch_sums"], expected) if check_sum else None
    fast_mode = engine.fast_mode
    row_pitch = engine.layout.row_pitch
    launches_per_epoch = None
    if timed_ep:
        launches_per_epoch = res["scatter_launches"] / timed_ep
    engine_opts = {"exchange": args.exchange, "wait_mode": engine.wait_mode,
                   "backpressure": engine.backpressure, "peer_alloc": engine.peer_alloc,
                   "tmap_mode": engine.tmap_mode, "sched": engine.sched,
                   "row_bytes": row_pitch, "row_align": args.row_align,
                   "fast_mode": fast_mode, "chunk_passes": engine.chunk_passes,
                   "shuffle_priority": engine.shuffle_priority}
    ds.dataset.close()
    steps = res["steps"]
    rows = timed_ep * args.rows_per_gpu * world       # rows delivered inside the region
    value = rows / (res["ms"] / 1e3)

    # ---- end to end: pinned host table, H2D every step, D2H every step ---------
    e2e = None
    if not args.skip_e2e:
        ds2 = make_dataset(args, files, rank, world, epochs, "host", torch)
        eng2 = ds2.dataset.engine
        res2 = run_phase(ds2, eng2, torch, dist, world, warm_ep, timed_ep, d2h_each_step=True)
        once2 = exactly_once(res2["epoch_sums"], expected) if check_sum else None
        ds2.dataset.close()
        e2e = {"value": rows / res2["wall"], "unit": "rows/s",
               # counted from the cudaMemcpyAsync calls enqueued inside the region
               "h2d_bytes_per_step": int(res2["h2d_bytes"] / max(1, res2["steps"])),
               "d2h_bytes_per_step": 8, "ms_per_step": res2["wall"] * 1e3 / res2["steps"],
               "device_ms_per_step": res2["ms"] / res2["steps"], "steps": res2["steps"],
               "gpu_launches": res2["launches"],
               "scatter_launches_in_region": res2["scatter_launches"],
               "h2d_bytes_in_region": res2["h2d_bytes"],
               "h2d_gbps_per_gpu": res2["h2d_bytes"] / res2["wall"] / 1e9,
               "batches_per_sec": rows / res2["wall"] / args.batch_size,
               "exactly_once": once2}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # bytes the shuffle moves per epoch per GPU: read source once + write rows once
        epoch_bytes = args.rows_per_gpu * (src_row_bytes + row_pitch)
        kms = sorted(res["kernel_ms"])
        # median over the timed epochs, of the SLOWEST rank (that one sets the epoch time)
        kernel_ms = max(res["kernel_ms_per_rank"]) or None
        egress = (args.rows_per_gpu * row_pitch * (world - 1) / world / (kernel_ms / 1e3) / 1e9
                  if kernel_ms and world > 1 else None)
        out = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": steps, "warmup": res["warm_steps"],
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "epochs_timed": timed_ep, "epochs_warmup": warm_ep,
            "batches_per_epoch": batches_per_epoch,
            "ms_per_step": res["ms"] / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.feature_dtype, "data": "synthetic",
            "impl": "ours",
            "config": shape_config(args, world),
            "engine": engine_opts,
            "batches_per_sec": value / args.batch_size,
            "e2e": e2e, "gpu_launches": res["launches"], "clocks": clocks,
            "scatter_launches_in_region": res["scatter_launches"],
            "scatter_launches_per_epoch": launches_per_epoch,
            "wall_ms_per_step": res["wall"] * 1e3 / steps,
            "ms_per_epoch": res["ms"] / timed_ep,
            # host time spent enqueueing the region's steps (Python iterator + launches);
            # well below ms_per_epoch = the device, not the host, sets the pace
            "host_enqueue_ms_per_epoch": res["host_loop_s"] * 1e3 / timed_ep,
            "shuffle_kernel_ms_per_epoch": kernel_ms,
            # K7: device time from the start of an epoch's shuffle until the first
            # reducer chunk is consumable (== the whole epoch with a single pass)
            "first_chunk_ms": res["first_pass_ms"] or kernel_ms,
            "shuffle_kernel_ms_min_max": [kms[0], kms[-1]] if kms else None,
            "shuffle_kernel_ms_per_rank": res["kernel_ms_per_rank"],
            "shuffle_kernel_gbps": (epoch_bytes / (kernel_ms / 1e3) / 1e9 if kernel_ms else None),
            "hbm_roofline_frac_of_measured": (
                epoch_bytes / (kernel_ms / 1e3) / 1e9 / peaks["hbm_gbs"]
                if kernel_ms and peaks.get("hbm_gbs") else None),
            "nvlink_egress_gbps_per_gpu": egress,
            "nvlink_frac_of_900": egress / 900.0 if egress else None,
            "nvlink_frac_of_measured_peer_copy_774": egress / 774.0 if egress else None,
            "exactly_once": once,
            "ingest_seconds": ingest_s,
            "cold_rows_per_sec": (args.rows_per_gpu * world / ingest_s if ingest_s else None),
            "datagen_seconds": gen_s,
        }
        if once is not None and not once["ok"]:
            out["invalid"] = "exactly-once check failed"
        if res["scatter_launches"] < timed_ep:
            out["invalid"] = "no shuffle kernel launched inside the timed region"
        emit_json(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not args.keep_data and rank == 0:
        d, _ = dataset_files(args, world)
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------

def run_reference(args):
    from baseline import reference_arm
    reference_arm.main(args)


if __name__ == "__main__":
    a = parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.min_timed_epochs is None:
        a.min_timed_epochs = 20 if a.impl == "ours" else 1
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
