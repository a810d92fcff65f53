"""Shuffle benchmark harness (component C15): mock trainers, trials, stats CSVs.

Same CLI surface and metric definitions as the reference's
``benchmarks/benchmark.py`` (flags ``:188-215``; ``row_throughput =
num_epochs * num_rows / duration`` etc., ``stats.py:396-401``): it times the
public ``shuffle()`` driver feeding mock consumers that only count rows, runs
``--num-trials`` (or ``--trials-timeout`` seconds of) trials, samples store
utilisation, prints the summary and writes the trial/epoch/consumer CSVs.

Differences: consumers are in-process objects with their own epoch window (no
Ray actors / placement groups); ``--cluster`` means "join the ambient
torch.distributed job" (launch with torchrun, one rank per GPU);
``--object-store-memory`` is accepted and ignored; new flags select the backend,
the schema (``--num-columns`` float32 columns instead of DATA_SPEC), the seed
and the exchange (fused P2P scatter vs the NCCL baseline).
"""
from __future__ import annotations

import argparse
import collections
import contextlib
import glob
import os
import sys
import threading
import timeit

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ray_shuffling_data_loader_b200.shuffle import shuffle, BatchConsumer as _BatchConsumer  # noqa: E402
from ray_shuffling_data_loader_b200.stats import (  # noqa: E402
    TrialStatsCollector, ObjectStoreStatsCollector, process_stats, human_readable_size)
from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec  # noqa: E402

DEFAULT_DATA_DIR = "/tmp/benchmark_scratch"
DEFAULT_STATS_DIR = "./results"
DEFAULT_UTILIZATION_SAMPLE_PERIOD = 5.0


class Consumer:
    """Mock trainer: counts rows, enforces its own epoch window
    (reference benchmark.py:29-62)."""

    def __init__(self, rank, num_epochs, max_concurrent_epochs, stats_collector=None,
                 verbose=True):
        self._rank = rank
        self._num_epochs = num_epochs
        self._max_epochs = max_concurrent_epochs
        self._curr_epochs = collections.deque()
        self._epoch_done_evs = [threading.Event() for _ in range(num_epochs)]
        self._stats_collector = stats_collector
        self._verbose = verbose
        self.rows = 0

    def new_epoch(self, epoch):
        if len(self._curr_epochs) == self._max_epochs:
            first_epoch = self._curr_epochs.popleft()
            self._epoch_done_evs[first_epoch].wait()
        self._curr_epochs.append(epoch)
        if self._verbose:
            print(f"Starting epoch {epoch} on consumer {self._rank}.")

    def consume(self, epoch, batch):
        batch.wait()                 # the data really has to arrive
        self.rows += len(batch)
        if self._stats_collector is not None:
            self._stats_collector.consume_batch(epoch, len(batch))

    def producer_done(self, epoch):
        if self._stats_collector is not None:
            self._stats_collector.consume_done(epoch)
        self._epoch_done_evs[epoch].set()
        if self._verbose:
            print(f"Epoch {epoch} done on consumer {self._rank}.")

    def wait_until_all_epochs_done(self):
        self._epoch_done_evs[self._num_epochs - 1].wait()

    def reset(self):
        self._curr_epochs.clear()
        for ev in self._epoch_done_evs:
            ev.clear()


class BatchConsumer(_BatchConsumer):
    def __init__(self, num_trainers, num_epochs, max_concurrent_epochs,
                 stats_collector=None, verbose=True):
        self._consumers = [Consumer(rank, num_epochs, max_concurrent_epochs,
                                    stats_collector, verbose)
                           for rank in range(num_trainers)]
        self._buffers = {}

    def consume(self, rank, epoch, batches):
        if batches is not None:
            for batch in batches:
                self._consumers[rank].consume(epoch, batch)
                self._buffers[(epoch, rank)] = batch.buffer

    def producer_done(self, rank, epoch):
        buf = self._buffers.pop((epoch, rank), None)
        if buf is not None:
            buf.release()            # frees the epoch-ring slot for reuse
        self._consumers[rank].producer_done(epoch)

    def wait_until_ready(self, epoch):
        for consumer in self._consumers:
            consumer.new_epoch(epoch)

    def wait_until_all_epochs_done(self):
        for consumer in self._consumers:
            consumer.wait_until_all_epochs_done()

    def reset(self):
        for c in self._consumers:
            c.reset()

    def rows(self):
        return sum(c.rows for c in self._consumers)


def run_trials(num_epochs, filenames, num_reducers, num_trainers, max_concurrent_epochs,
               utilization_sample_period, collect_stats=True, num_trials=None,
               trials_timeout=None, verbose=True, **engine_options):
    """
    Run shuffling trials.
    """
    print("Using from-memory shuffler.")
    all_stats = []
    if collect_stats:
        stats_collector = TrialStatsCollector(num_epochs, len(filenames), num_reducers,
                                              num_trainers)
        store_collector = ObjectStoreStatsCollector(utilization_sample_period)
    else:
        stats_collector = None
        store_collector = contextlib.nullcontext()
    batch_consumer = BatchConsumer(num_trainers, num_epochs, max_concurrent_epochs,
                                   stats_collector, verbose)

    def one_trial(trial):
        print(f"Starting trial {trial}.")
        batch_consumer.reset()
        if stats_collector is not None:
            stats_collector.reset()
        with store_collector:
            duration = shuffle(filenames, batch_consumer, num_epochs, num_reducers,
                               num_trainers, stats_collector,
                               max_concurrent_epochs=max_concurrent_epochs, **engine_options)
        print(f"Trial {trial} done after {duration} seconds.")
        if collect_stats:
            all_stats.append((stats_collector.get_stats(timeout=60),
                              store_collector.get_stats()))
        else:
            all_stats.append((duration, None))

    if num_trials is not None:
        for trial in range(num_trials):
            one_trial(trial)
    elif trials_timeout is not None:
        start = timeit.default_timer()
        trial = 0
        while timeit.default_timer() - start < trials_timeout:
            one_trial(trial)
            trial += 1
    else:
        raise ValueError("One of num_trials and trials_timeout must be specified")
    return all_stats


def build_parser():
    parser = argparse.ArgumentParser(description="Shuffling data loader")
    parser.add_argument("--num-rows", type=int, default=4 * (10**11))
    parser.add_argument("--num-files", type=int, default=100)
    parser.add_argument("--max-row-group-skew", type=float, default=0.0)
    parser.add_argument("--num-row-groups-per-file", type=int, default=1)
    parser.add_argument("--num-reducers", type=int, default=5)
    parser.add_argument("--num-trainers", type=int, default=5)
    parser.add_argument("--num-epochs", type=int, default=10)
    parser.add_argument("--max-concurrent-epochs", type=int, default=None)
    parser.add_argument("--batch-size", type=int, default=100)
    parser.add_argument("--num-trials", type=int, default=None)
    parser.add_argument("--trials-timeout", type=int, default=None)
    parser.add_argument("--utilization-sample-period", type=float,
                        default=DEFAULT_UTILIZATION_SAMPLE_PERIOD)
    parser.add_argument("--cluster", action="store_true")
    parser.add_argument("--object-store-memory", type=int, default=None)
    parser.add_argument("--data-dir", type=str, default=DEFAULT_DATA_DIR)
    parser.add_argument("--stats-dir", type=str, default=DEFAULT_STATS_DIR)
    parser.add_argument("--clear-old-data", action="store_true")
    parser.add_argument("--use-old-data", action="store_true")
    parser.add_argument("--no-stats", action="store_true")
    parser.add_argument("--no-epoch-stats", action="store_true")
    parser.add_argument("--no-consumer-stats", action="store_true")
    parser.add_argument("--overwrite-stats", action="store_true")
    parser.add_argument("--unique-stats", action="store_true")
    # B200 framework extensions
    parser.add_argument("--backend", choices=["auto", "cpu", "cuda"], default="auto")
    parser.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p")
    parser.add_argument("--resident", choices=["hbm", "host"], default="hbm")
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--num-columns", type=int, default=None,
                        help="generate N float32 columns instead of DATA_SPEC "
                             "(BASELINE.json wide-row sweep: 64/256/1024/4096)")
    parser.add_argument("--quiet", action="store_true")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)

    if args.num_row_groups_per_file < 1:
        raise ValueError("Must have at least one row group per file.")
    num_trials, trials_timeout = args.num_trials, args.trials_timeout
    if num_trials is not None and trials_timeout is not None:
        raise ValueError("Only one of --num-trials and --trials-timeout should be "
                         "specified.")
    if num_trials is None and trials_timeout is None:
        num_trials = 3
    if args.clear_old_data and args.use_old_data:
        raise ValueError("Only one of --clear-old-data and --use-old-data should be "
                         "specified.")

    data_dir = args.data_dir
    if args.clear_old_data:
        print(f"Clearing old data from {data_dir}.")
        for f in glob.glob(os.path.join(data_dir, "*.parquet.snappy")):
            os.remove(f)

    if args.cluster:
        if args.object_store_memory is not None:
            raise ValueError("Can't specify --object-store-memory when "
                             "connecting to existing cluster.")
        from ray_shuffling_data_loader_b200.parallel import bootstrap
        ctx = bootstrap.init_from_env()
        print(f"Joined a torch.distributed job: rank {ctx.rank} of {ctx.world}.")
    elif args.object_store_memory is not None:
        print(f"(--object-store-memory {human_readable_size(args.object_store_memory)} "
              "has no effect: HBM arenas are sized from the data)")

    num_rows, num_files = args.num_rows, args.num_files
    num_row_groups_per_file = args.num_row_groups_per_file
    max_row_group_skew = args.max_row_group_skew
    if not args.use_old_data:
        print(f"Generating {num_rows} rows over {num_files} files, with "
              f"{num_row_groups_per_file} row groups per file and at most "
              f"{100 * max_row_group_skew:.1f}% row group skew.")
        spec = float_spec(args.num_columns, np.float32) if args.num_columns else None
        filenames, num_bytes = generate_data(num_rows, num_files, num_row_groups_per_file,
                                             max_row_group_skew, data_dir, data_spec=spec,
                                             seed=args.seed)
        print(f"Generated {len(filenames)} files containing {num_rows} rows "
              f"with {num_row_groups_per_file} row groups per file, totalling "
              f"{human_readable_size(num_bytes)}.")
    else:
        filenames = [os.path.join(data_dir, f"input_data_{file_index}.parquet.snappy")
                     for file_index in range(num_files)]
        print("Not generating input data, using existing data instead.")

    num_reducers, num_trainers = args.num_reducers, args.num_trainers
    batch_size, num_epochs = args.batch_size, args.num_epochs
    max_concurrent_epochs = args.max_concurrent_epochs
    if max_concurrent_epochs is None or max_concurrent_epochs > num_epochs:
        max_concurrent_epochs = num_epochs
    assert max_concurrent_epochs > 0

    # TODO: warm-up trials (also a TODO upstream, benchmark.py:291).
    print("\nRunning real trials.")
    if num_trials is not None:
        print(f"Running {num_trials} shuffle trials with {num_epochs} epochs, "
              f"{num_reducers} reducers, {num_trainers} trainers, and a batch "
              f"size of {batch_size} over {num_rows} rows.")
    else:
        print(f"Running {trials_timeout} seconds of shuffle trials with "
              f"{num_epochs} epochs, {num_reducers} reducers, {num_trainers} "
              f"trainers, and a batch size of {batch_size} over {num_rows} rows.")
    print(f"Shuffling will be pipelined with at most "
          f"{max_concurrent_epochs} concurrent epochs.")
    collect_stats = not args.no_stats
    opts = {}
    backend = None if args.backend == "auto" else args.backend
    from ray_shuffling_data_loader_b200.runtime.engine import resolve_backend
    if resolve_backend(backend) == "cuda":
        opts.update(exchange=args.exchange, resident=args.resident)
    all_stats = run_trials(num_epochs, filenames, num_reducers, num_trainers,
                           max_concurrent_epochs, args.utilization_sample_period,
                           collect_stats, num_trials, trials_timeout,
                           verbose=not args.quiet, backend=backend, seed=args.seed,
                           batch_size=batch_size, **opts)

    if collect_stats:
        process_stats(all_stats, args.overwrite_stats, args.stats_dir,
                      args.no_epoch_stats, args.no_consumer_stats,
                      args.unique_stats, num_rows, num_files,
                      num_row_groups_per_file, batch_size, num_reducers,
                      num_trainers, num_epochs, max_concurrent_epochs)
    else:
        print("Shuffle trials done, no detailed stats collected.")
        times, _ = zip(*all_stats)
        mean, std = np.mean(times), np.std(times)
        throughput_std = np.std([num_epochs * num_rows / time for time in times])
        batch_throughput_std = np.std(
            [(num_epochs * num_rows / batch_size) / time for time in times])
        print(f"\nMean over {len(times)} trials: {mean:.3f}s +- {std}")
        print(f"Mean throughput over {len(times)} trials: "
              f"{num_epochs * num_rows / mean:.2f} rows/s +- {throughput_std:.2f}")
        print(f"Mean batch throughput over {len(times)} trials: "
              f"{(num_epochs * num_rows / batch_size) / mean:.2f} batches/s "
              f"+- {batch_throughput_std:.2f}")
    return all_stats


if __name__ == "__main__":
    main()
