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
    p.add_argument("--data-dir", default=os.environ.get("RSDL_BENCH_DIR",
                                                         os.path.join(tempfile.gettempdir(), "rsdl_bench")))
    p.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p")
    p.add_argument("--feature-dtype", choices=["float32", "bfloat16", "fp8"], default="float32")
    p.add_argument("--peer-alloc", choices=["symm", "ipc"], default=None)
    p.add_argument("--wait-mode", choices=["stream", "host"], default="stream",
                   help="how a trainer waits for an epoch's produced flags: a wait kernel on its "
                        "own stream (no host round trip) or a host poll")
    p.add_argument("--max-concurrent-epochs", type=int, default=2,
                   help="epoch window (BASELINE config 3 compares 1 vs 2)")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--keep-data", action="store_true")
    p.add_argument("--ref-steps-cap", type=int, default=None,
                   help="reference arm: cap on timed steps (it is slow)")
    return p.parse_args()


# ---------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------

def dataset_files(args, world):
    n_files = FILES_PER_GPU * world
    d = os.path.join(args.data_dir, f"r{args.rows_per_gpu}_c{args.cols}_w{world}")
    return d, [os.path.join(d, f"input_data_{i}.parquet.snappy") for i in range(n_files)]


def generate_my_share(args, rank, world):
    """Every rank writes FILES_PER_GPU files of the global table (outside any
    timed region). Global row index = position in the concatenation."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from ray_shuffling_data_loader_b200.data_generation import float_spec, generate_file
    d, files = dataset_files(args, world)
    os.makedirs(d, exist_ok=True)
    spec = float_spec(args.cols, np.float32)
    rows_per_file = args.rows_per_gpu // FILES_PER_GPU
    mine = range(rank * FILES_PER_GPU, (rank + 1) * FILES_PER_GPU)
    todo = [i for i in mine if not os.path.exists(files[i] + ".ok")]

    def one(i):
        rows = rows_per_file + (args.rows_per_gpu - rows_per_file * FILES_PER_GPU
                                if i % FILES_PER_GPU == FILES_PER_GPU - 1 else 0)
        start = (i // FILES_PER_GPU) * args.rows_per_gpu + (i % FILES_PER_GPU) * rows_per_file
        generate_file(i, start, rows, ROW_GROUPS_PER_FILE, d, spec,
                      np.random.SeedSequence([1234, i]), include_key=False)
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
# our arm
# ---------------------------------------------------------------------------

def run_phase(ds, engine, torch, dist, world, steps, warmup, batch_size, d2h_each_step):
    """Consume ``warmup`` then ``steps`` batches. The clock starts BEFORE the first
    timed batch is fetched, so waiting for that batch's epoch to be shuffled is
    inside the timed region (no pre-shuffled epoch is consumed for free).
    Returns (max-over-ranks device ms, wall s, launches in the region, checksum)."""
    row_pitch = engine.layout.row_pitch
    dev = torch.device("cuda", torch.cuda.current_device())
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    host_acc = torch.zeros(1, dtype=torch.float64).pin_memory()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    state = {"it": None, "epoch": 0}

    def next_batch():
        while True:
            if state["it"] is None:
                ds.set_epoch(state["epoch"])
                state["it"] = iter(ds)
            try:
                return next(state["it"])       # the public API: (features, label)
            except StopIteration:
                state["it"] = None
                state["epoch"] += 1

    checksum = [0.0]

    def step():
        features, label = next_batch()
        # consume: every byte of the batch (features + label share one packed row;
        # the feature view starts at the batch's first byte) is read by our kernel
        base = features[0] if isinstance(features, tuple) else features
        engine.batch_sum_all(base, acc, nbytes=base.shape[0] * row_pitch)
        if d2h_each_step:
            host_acc.copy_(acc, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            checksum[0] = float(host_acc[0])

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = engine.launches
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = engine.launches - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    wall_t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
    if not d2h_each_step:
        checksum[0] = float(acc.item())
    return float(ms.item()), float(wall_t.item()), launches, checksum[0]


def make_dataset(args, files, rank, world, epochs, resident, torch, seed=20260921):
    """The flagship public API, exactly as a user would call it."""
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    feature_columns = [f"f{i}" for i in range(args.cols - 1)]
    dt = {"float32": torch.float32, "bfloat16": torch.bfloat16,
          "fp8": getattr(torch, "float8_e4m3fn", None)}[args.feature_dtype]
    opts = dict(resident=resident, exchange=args.exchange, wait_mode=args.wait_mode)
    if args.peer_alloc:
        opts["peer_alloc"] = args.peer_alloc
    if resident == "host":
        opts["stream_chunk_rows"] = args.batch_size
    fp8 = args.feature_dtype == "fp8"
    return TorchShufflingDataset(
        files, epochs, world, args.batch_size, rank, num_reducers=world,
        max_concurrent_epochs=args.max_concurrent_epochs, feature_columns=feature_columns,
        feature_types=[dt] * len(feature_columns), label_column="labels",
        label_type=dt if not fp8 else torch.float32, packed_features=True,
        fp8_block_scale=fp8, seed=seed, backend="cuda", queue_name=f"bench-{resident}", **opts)


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
    files_mine = generate_my_share(args, rank, world)
    if world > 1:
        dist.barrier()
    gen_s = time.perf_counter() - t_gen
    _, files = dataset_files(args, world)
    batches_per_epoch = -(-args.rows_per_gpu // args.batch_size)
    total_steps = args.steps + args.warmup
    epochs = -(-total_steps // batches_per_epoch) + 1
    sampler = ClockSampler(range(world)) if rank == 0 else None

    # ---- device-timed, HBM-resident -----------------------------------------
    t0 = time.perf_counter()
    ds = make_dataset(args, files, rank, world, epochs, "hbm", torch)
    engine = ds.dataset.engine
    if sampler:
        sampler.start()
    ms, wall, launches, chk = run_phase(ds, engine, torch, dist, world, args.steps, args.warmup,
                                        args.batch_size, d2h_each_step=False)
    clocks = sampler.stop() if sampler else None
    ingest_s = getattr(engine, "ingest_seconds", None)
    kernel_ms = [engine.epoch_kernel_ms(e) for e in range(epochs) if engine.epoch_kernel_ms(e)]
    fast_mode = engine.fast_mode
    row_pitch = engine.layout.row_pitch
    ds.dataset.close()
    rows = args.steps * args.batch_size * world
    value = rows / (ms / 1e3)

    # ---- end to end: pinned host table, H2D every step, D2H every step ---------
    e2e = None
    if not args.skip_e2e:
        ds2 = make_dataset(args, files, rank, world, epochs, "host", torch)
        eng2 = ds2.dataset.engine
        ms2, wall2, launches2, chk2 = run_phase(ds2, eng2, torch, dist, world, args.steps,
                                                args.warmup, args.batch_size, d2h_each_step=True)
        h2d_epoch = eng2.h2d_bytes_per_epoch()
        ds2.dataset.close()
        e2e = {"value": rows / wall2, "unit": "rows/s",
               "h2d_bytes_per_step": int(h2d_epoch / batches_per_epoch),
               "d2h_bytes_per_step": 8, "ms_per_step": wall2 * 1e3 / args.steps,
               "device_ms_per_step": ms2 / args.steps, "gpu_launches": launches2,
               "batches_per_sec": rows / wall2 / args.batch_size}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # bytes the shuffle moves per epoch per GPU: read source once + write rows once
        epoch_bytes = args.rows_per_gpu * (args.cols * 4 + row_pitch)
        best_kernel_ms = min(kernel_ms) if kernel_ms else None
        out = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.feature_dtype, "data": "synthetic",
            "impl": "ours",
            "config": {"model": f"TorchShufflingDataset {world} trainers x {world} reducers",
                       "global_batch": args.batch_size * world,
                       "rows": args.rows_per_gpu * world, "cols": args.cols,
                       "row_bytes": row_pitch, "batch_size": args.batch_size,
                       "seq_len": None, "parallelism": f"dp{world}",
                       "max_concurrent_epochs": args.max_concurrent_epochs, "exchange": args.exchange,
                       "wait_mode": args.wait_mode,
                       "l2_policy": "inputs larger than L2 (3.2 GB/epoch/GPU)"},
            "batches_per_sec": value / args.batch_size,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "wall_ms_per_step": wall * 1e3 / args.steps,
            "shuffle_kernel_ms_per_epoch": best_kernel_ms,
            "shuffle_kernel_gbps": (epoch_bytes / (best_kernel_ms / 1e3) / 1e9
                                    if best_kernel_ms else None),
            "hbm_roofline_frac_of_measured": (
                epoch_bytes / (best_kernel_ms / 1e3) / 1e9 / peaks["hbm_gbs"]
                if best_kernel_ms and peaks.get("hbm_gbs") else None),
            "nvlink_egress_gbps_per_gpu": (
                args.rows_per_gpu * row_pitch * (world - 1) / world / (best_kernel_ms / 1e3) / 1e9
                if best_kernel_ms and world > 1 else None),
            "ingest_seconds": ingest_s, "datagen_seconds": gen_s, "fast_mode": fast_mode,
            "checksum": chk,
        }
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
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
