"""Shuffle benchmark harness (component C15): mock trainers, trials, stats CSVs.

Same command line and metric definitions as the reference's
``benchmarks/benchmark.py`` (its flags ``:188-215``; ``row_throughput =
num_epochs * num_rows / duration`` etc., ``stats.py:396-401``) so existing sweep
scripts keep working: it times the public ``shuffle()`` driver feeding mock
trainers that only count rows, runs ``--num-trials`` (or ``--trials-timeout``
seconds of) trials, samples store utilisation, prints the summary and writes the
trial / epoch / consumer CSVs.

What is different underneath: the mock trainers are one in-process object
(``MockTrainers``) that keeps a single epoch window for all ranks instead of one
Ray actor per rank in a placement group; a delivered chunk is a handle to rows that
are already in the trainer's HBM, so "consuming" it is waiting for its completion
flag; ``--cluster`` means "join the ambient torch.distributed job" (launch with
torchrun, one rank per GPU); ``--object-store-memory`` is accepted and ignored.
Extra flags select the backend, the schema (``--num-columns`` float32 columns
instead of ``DATA_SPEC``), the seed, the residency mode and the exchange (fused P2P
scatter vs the NCCL baseline).
"""
from __future__ import annotations

import argparse
import contextlib
import glob
import os
import sys
import threading
import timeit
from typing import Dict, List, Optional

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ray_shuffling_data_loader_b200.shuffle import shuffle, BatchConsumer  # noqa: E402
from ray_shuffling_data_loader_b200.stats import (  # noqa: E402
    TrialStatsCollector, ObjectStoreStatsCollector, process_stats, human_readable_size)
from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec  # noqa: E402

# (flag, kwargs) - the reference's CLI, flag for flag (benchmark.py:188-215) ...
REFERENCE_FLAGS = [
    ("--num-rows", dict(type=int, default=4 * (10**11))),
    ("--num-files", dict(type=int, default=100)),
    ("--max-row-group-skew", dict(type=float, default=0.0)),
    ("--num-row-groups-per-file", dict(type=int, default=1)),
    ("--num-reducers", dict(type=int, default=5)),
    ("--num-trainers", dict(type=int, default=5)),
    ("--num-epochs", dict(type=int, default=10)),
    ("--max-concurrent-epochs", dict(type=int, default=None)),
    ("--batch-size", dict(type=int, default=100)),
    ("--num-trials", dict(type=int, default=None)),
    ("--trials-timeout", dict(type=int, default=None)),
    ("--utilization-sample-period", dict(type=float, default=5.0)),
    ("--cluster", dict(action="store_true")),
    ("--object-store-memory", dict(type=int, default=None)),
    ("--data-dir", dict(type=str, default="/tmp/benchmark_scratch")),
    ("--stats-dir", dict(type=str, default="./results")),
    ("--clear-old-data", dict(action="store_true")),
    ("--use-old-data", dict(action="store_true")),
    ("--no-stats", dict(action="store_true")),
    ("--no-epoch-stats", dict(action="store_true")),
    ("--no-consumer-stats", dict(action="store_true")),
    ("--overwrite-stats", dict(action="store_true")),
    ("--unique-stats", dict(action="store_true")),
]
# ... plus what only exists here
EXTENSION_FLAGS = [
    ("--backend", dict(choices=["auto", "cpu", "cuda"], default="auto")),
    ("--exchange", dict(choices=["p2p", "nccl"], default="p2p")),
    ("--resident", dict(choices=["hbm", "host", "disk"], default="hbm")),
    ("--chunk-passes", dict(type=int, default=None,
                            help="destination-chunk passes per epoch (K7; default: engine's)")),
    ("--seed", dict(type=int, default=None)),
    ("--num-columns", dict(type=int, default=None,
                           help="generate N float32 columns instead of DATA_SPEC "
                                "(BASELINE.json wide-row sweep: 64/256/1024/4096)")),
    ("--quiet", dict(action="store_true")),
]


class MockTrainers(BatchConsumer):
    """All mock trainers of a trial (the role of the reference's ``Consumer`` actors
    plus its fan-out ``BatchConsumer``, benchmark.py:29-108).

    A trainer "consumes" a chunk by waiting for its completion flag and counting its
    rows; an epoch is finished for a rank at ``producer_done``. The epoch window -
    at most ``window`` epochs between "shuffle may start" and "every rank finished" -
    is one counter under a condition variable."""

    def __init__(self, num_trainers: int, num_epochs: int, window: int,
                 stats: Optional[TrialStatsCollector] = None, verbose: bool = True):
        self.num_trainers, self.num_epochs, self.window = num_trainers, num_epochs, window
        self.stats, self.verbose = stats, verbose
        self._cv = threading.Condition()
        self.reset()

    def reset(self):
        with self._cv:
            self._ranks_done: Dict[int, int] = {}     # epoch -> ranks that finished it
            self._admitted: List[int] = []            # epochs whose shuffle was let through
            self._held = {}                           # (epoch, rank) -> epoch buffer to release
            self.rows = 0

    def _finished(self, epoch: int) -> bool:
        return self._ranks_done.get(epoch, 0) >= self.num_trainers

    # -- BatchConsumer ---------------------------------------------------------
    def wait_until_ready(self, epoch):
        with self._cv:
            self._cv.wait_for(lambda: sum(1 for e in self._admitted if not self._finished(e))
                              < self.window)
            self._admitted.append(epoch)
        if self.verbose:
            print(f"epoch {epoch}: shuffle admitted")

    def consume(self, rank, epoch, batches):
        rows = 0
        for chunk in batches or ():
            chunk.wait()                    # the rows really have to be there
            rows += len(chunk)
            if self.stats is not None:
                self.stats.consume_batch(epoch, len(chunk))
            self._held[(epoch, rank)] = chunk.buffer
        with self._cv:
            self.rows += rows

    def producer_done(self, rank, epoch):
        buf = self._held.pop((epoch, rank), None)
        if buf is not None:
            buf.release()                   # hands the epoch-ring slot back
        if self.stats is not None:
            self.stats.consume_done(epoch)
        with self._cv:
            self._ranks_done[epoch] = self._ranks_done.get(epoch, 0) + 1
            self._cv.notify_all()
        if self.verbose:
            print(f"epoch {epoch}: trainer {rank} done")

    def wait_until_all_epochs_done(self):
        with self._cv:
            self._cv.wait_for(lambda: self._finished(self.num_epochs - 1))


def run_trials(num_epochs, filenames, num_reducers, num_trainers, max_concurrent_epochs,
               utilization_sample_period, collect_stats=True, num_trials=None,
               trials_timeout=None, verbose=True, **engine_options):
    """Run ``num_trials`` trials (or trials until ``trials_timeout`` seconds have
    passed) of ``shuffle()`` -> ``[(TrialStats | duration, store stats | None), ...]``."""
    if (num_trials is None) == (trials_timeout is None):
        raise ValueError("One of num_trials and trials_timeout must be specified")
    stats = store = None
    if collect_stats:
        stats = TrialStatsCollector(num_epochs, len(filenames), num_reducers, num_trainers)
        store = ObjectStoreStatsCollector(utilization_sample_period)
    trainers = MockTrainers(num_trainers, num_epochs, max_concurrent_epochs, stats, verbose)
    results = []
    started = timeit.default_timer()
    trial = 0
    while (trial < num_trials) if num_trials is not None \
            else (timeit.default_timer() - started < trials_timeout):
        print(f"Starting trial {trial}.")
        trainers.reset()
        if stats is not None:
            stats.reset()
        with (store if store is not None else contextlib.nullcontext()):
            duration = shuffle(filenames, trainers, num_epochs, num_reducers, num_trainers,
                               stats, max_concurrent_epochs=max_concurrent_epochs,
                               **engine_options)
        print(f"Trial {trial} done after {duration} seconds.")
        results.append((stats.get_stats(timeout=60), store.get_stats()) if collect_stats
                       else (duration, None))
        trial += 1
    return results


def build_parser():
    parser = argparse.ArgumentParser(description="Shuffling data loader")
    for flag, kw in REFERENCE_FLAGS + EXTENSION_FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def _validate(args):
    if args.num_row_groups_per_file < 1:
        raise ValueError("Must have at least one row group per file.")
    if args.num_trials is not None and args.trials_timeout is not None:
        raise ValueError("Only one of --num-trials and --trials-timeout should be specified.")
    if args.clear_old_data and args.use_old_data:
        raise ValueError("Only one of --clear-old-data and --use-old-data should be specified.")
    if args.cluster and args.object_store_memory is not None:
        raise ValueError("Can't specify --object-store-memory when connecting to existing "
                         "cluster.")
    if args.num_trials is None and args.trials_timeout is None:
        args.num_trials = 3                      # the reference's default (benchmark.py:227-228)
    window = args.max_concurrent_epochs
    if window is None or window > args.num_epochs:
        window = args.num_epochs                 # reference benchmark.py:284-287
    if window < 1:
        raise ValueError("--max-concurrent-epochs must be positive")
    args.max_concurrent_epochs = window


def _input_files(args) -> List[str]:
    if args.clear_old_data:
        print(f"Clearing old data from {args.data_dir}.")
        for f in glob.glob(os.path.join(args.data_dir, "*.parquet.snappy")):
            os.remove(f)
    if args.use_old_data:
        print("Not generating input data, using existing data instead.")
        return [os.path.join(args.data_dir, f"input_data_{i}.parquet.snappy")
                for i in range(args.num_files)]
    print(f"Generating {args.num_rows} rows over {args.num_files} files, with "
          f"{args.num_row_groups_per_file} row groups per file and at most "
          f"{100 * args.max_row_group_skew:.1f}% row group skew.")
    spec = float_spec(args.num_columns, np.float32) if args.num_columns else None
    filenames, num_bytes = generate_data(args.num_rows, args.num_files,
                                         args.num_row_groups_per_file, args.max_row_group_skew,
                                         args.data_dir, data_spec=spec, seed=args.seed)
    print(f"Generated {len(filenames)} files containing {args.num_rows} rows with "
          f"{args.num_row_groups_per_file} row groups per file, totalling "
          f"{human_readable_size(num_bytes)}.")
    return filenames


def _print_timing_only(durations, args):
    """--no-stats: the three headline means with their spread over trials."""
    d = np.asarray(durations, dtype=np.float64)
    total_rows = args.num_epochs * args.num_rows
    rows_s, batches_s = total_rows / d, total_rows / args.batch_size / d
    n = len(d)
    print("Shuffle trials done, no detailed stats collected.")
    print(f"\nMean over {n} trials: {d.mean():.3f}s +- {d.std()}")
    print(f"Mean throughput over {n} trials: {total_rows / d.mean():.2f} rows/s "
          f"+- {rows_s.std():.2f}")
    print(f"Mean batch throughput over {n} trials: "
          f"{total_rows / args.batch_size / d.mean():.2f} batches/s +- {batches_s.std():.2f}")


def main(argv=None):
    args = build_parser().parse_args(argv)
    _validate(args)
    if args.cluster:
        from ray_shuffling_data_loader_b200.parallel import bootstrap
        ctx = bootstrap.init_from_env()
        print(f"Joined a torch.distributed job: rank {ctx.rank} of {ctx.world}.")
    elif args.object_store_memory is not None:
        print(f"(--object-store-memory {human_readable_size(args.object_store_memory)} "
              "has no effect: HBM arenas are sized from the data)")
    filenames = _input_files(args)

    what = (f"{args.num_trials} shuffle trials" if args.num_trials is not None
            else f"{args.trials_timeout} seconds of shuffle trials")
    print(f"\nRunning {what} with {args.num_epochs} epochs, {args.num_reducers} reducers, "
          f"{args.num_trainers} trainers, and a batch size of {args.batch_size} over "
          f"{args.num_rows} rows; at most {args.max_concurrent_epochs} concurrent epochs.")
    backend = None if args.backend == "auto" else args.backend
    engine_options = dict(backend=backend, seed=args.seed, batch_size=args.batch_size)
    from ray_shuffling_data_loader_b200.runtime.engine import resolve_backend
    if resolve_backend(backend) == "cuda":
        engine_options.update(exchange=args.exchange, resident=args.resident)
        if args.chunk_passes is not None:
            engine_options["chunk_passes"] = args.chunk_passes
    collect_stats = not args.no_stats
    all_stats = run_trials(args.num_epochs, filenames, args.num_reducers, args.num_trainers,
                           args.max_concurrent_epochs, args.utilization_sample_period,
                           collect_stats, args.num_trials, args.trials_timeout,
                           verbose=not args.quiet, **engine_options)
    if collect_stats:
        process_stats(all_stats, args.overwrite_stats, args.stats_dir, args.no_epoch_stats,
                      args.no_consumer_stats, args.unique_stats, args.num_rows, args.num_files,
                      args.num_row_groups_per_file, args.batch_size, args.num_reducers,
                      args.num_trainers, args.num_epochs, args.max_concurrent_epochs)
    else:
        _print_timing_only([d for d, _ in all_stats], args)
    return all_stats


if __name__ == "__main__":
    main()
