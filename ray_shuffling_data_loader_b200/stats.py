"""Shuffle statistics: collectors, utilisation sampler and CSV reports (C11-C13).

Same data model and the same three CSV schemas as the reference
(``ray_shuffling_data_loader/stats.py``: dataclasses ``:24-64``, collectors
``:68-255``, store sampler ``:258-279``, ``process_stats`` ``:287-625``,
``human_readable_*`` ``:631-646``), so existing analysis notebooks keep working.
The stage names keep the reference's vocabulary and map onto this framework as

    map     = ingest / staging of one source unit (Parquet decode + H2D)
    reduce  = production of one reducer chunk (the scatter kernel's share)
    consume = hand-off of one reducer chunk to a trainer

Differences, all forced by the substrate: collectors are plain thread-safe
objects called directly (no Ray actor, no ``.remote``); the object-store sampler
polls a byte-count callback (HBM arena bytes in use) instead of scraping the
raylet over gRPC; two optional trailing CSV columns report the measured
exchange bandwidth and its fraction of the NVLink roofline.
"""
from __future__ import annotations

import csv
import datetime
import math
import os
import threading
import timeit
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

try:  # fsspec lets stats go to s3:// like the reference; local files otherwise
    import fsspec
except Exception:  # pragma: no cover
    fsspec = None

#
# Stats data classes.
#


@dataclass
class StageStats:
    task_durations: List[float]
    stage_duration: float


@dataclass
class MapStats(StageStats):
    read_durations: List[float]


@dataclass
class ReduceStats(StageStats):
    pass


@dataclass
class ConsumeStats:
    stage_duration: float
    consume_times: Dict[float, int]
    time_to_consumes: List[float]


@dataclass
class ThrottleStats:
    wait_duration: float


@dataclass
class EpochStats:
    duration: float
    map_stats: MapStats
    reduce_stats: ReduceStats
    consume_stats: ConsumeStats
    throttle_stats: ThrottleStats
    # Device-side extras (0 when unknown): bytes pushed over the fabric and the
    # CUDA-event duration of the exchange kernel(s), max over ranks.
    exchange_bytes: int = 0
    exchange_seconds: float = 0.0


@dataclass
class TrialStats:
    epoch_stats: List[EpochStats]
    duration: float


@dataclass
class StoreSample:
    """One utilisation sample; attribute name kept from the raylet reply."""
    object_store_bytes_used: int


#
# Shuffling data loader stats collectors.
#


class _Stage:
    """Start/finish bookkeeping for one stage of one epoch."""

    def __init__(self, expected: int):
        self.expected = expected
        self.started = 0
        self.finished = 0
        self.t_first: Optional[float] = None
        self.duration: Optional[float] = None
        self.task_durations: List[float] = []

    def begin(self, now: float):
        if self.started == 0:
            self.t_first = now
        self.started += 1

    def end(self, now: float, task_duration: Optional[float] = None) -> bool:
        self.finished += 1
        if task_duration is not None:
            self.task_durations.append(task_duration)
        if self.finished == self.expected:
            if self.t_first is None:
                self.t_first = now
            self.duration = now - self.t_first
            return True
        return False


class EpochStatsCollector_:
    def __init__(self, num_maps, num_reduces, num_consumes):
        self._map = _Stage(num_maps)
        self._reduce = _Stage(num_reduces)
        self._consume = _Stage(num_consumes)
        self._num_reduces = num_reduces
        self._read_durations: List[float] = []
        self._consume_times: Dict[float, int] = {}
        self._time_to_consumes: List[float] = []
        self._t0: Optional[float] = None
        self._duration: Optional[float] = None
        self._throttle: Optional[float] = None
        self._exchange_bytes = 0
        self._exchange_seconds = 0.0
        self._done = threading.Event()

    def epoch_start(self):
        self._t0 = timeit.default_timer()
        self._consume_times[self._t0] = 0

    def map_start(self):
        self._map.begin(timeit.default_timer())

    def map_done(self, duration, read_duration):
        self._read_durations.append(read_duration)
        self._map.end(timeit.default_timer(), duration)

    def reduce_start(self):
        self._reduce.begin(timeit.default_timer())

    def reduce_done(self, duration):
        now = timeit.default_timer()
        if self._reduce.end(now, duration):
            # The epoch is "done" when its last reducer is (reference :155-158).
            assert self._t0 is not None
            self._duration = now - self._t0
            self._done.set()

    def consume_batch(self, num_rows):
        now = timeit.default_timer()
        if self._consume.t_first is None:
            self._consume.t_first = now
        self._consume_times[now] = num_rows
        self._time_to_consumes.append(now - (self._t0 if self._t0 is not None else now))

    def consume_done(self):
        self._consume.end(timeit.default_timer())

    def throttle_done(self, duration):
        self._throttle = duration

    def exchange_done(self, nbytes: int, seconds: float):
        self._exchange_bytes += int(nbytes)
        self._exchange_seconds = max(self._exchange_seconds, float(seconds))

    def get_stats(self, timeout: Optional[float] = None) -> EpochStats:
        if not self._done.wait(timeout):
            raise TimeoutError("epoch did not finish")
        consume_duration = self._consume.duration
        if consume_duration is None:
            consume_duration = 0.0
        return EpochStats(
            self._duration,
            MapStats(self._map.task_durations, self._map.duration or 0.0,
                     self._read_durations),
            ReduceStats(self._reduce.task_durations, self._reduce.duration or 0.0),
            ConsumeStats(consume_duration, self._consume_times, self._time_to_consumes),
            ThrottleStats(self._throttle or 0),
            self._exchange_bytes, self._exchange_seconds)


class TrialStatsCollector_:
    """Fans events out to per-epoch collectors. Thread-safe; reusable across
    trials (``reset``)."""

    def __init__(self, num_epochs, num_maps, num_reduces, num_consumes):
        self._shape = (num_epochs, num_maps, num_reduces, num_consumes)
        self._lock = threading.Lock()
        self.reset()

    def reset(self):
        num_epochs, num_maps, num_reduces, num_consumes = self._shape
        self._collectors = [EpochStatsCollector_(num_maps, num_reduces, num_consumes)
                            for _ in range(num_epochs)]
        self._duration = None
        self._trial_done_ev = threading.Event()

    def trial_done(self, duration):
        self._duration = duration
        self._trial_done_ev.set()

    def epoch_throttle_done(self, epoch, duration):
        with self._lock:
            self._collectors[epoch].throttle_done(duration)

    def epoch_start(self, epoch):
        with self._lock:
            self._collectors[epoch].epoch_start()

    def map_start(self, epoch):
        with self._lock:
            self._collectors[epoch].map_start()

    def map_done(self, epoch, duration, read_duration):
        with self._lock:
            self._collectors[epoch].map_done(duration, read_duration)

    def reduce_start(self, epoch):
        with self._lock:
            self._collectors[epoch].reduce_start()

    def reduce_done(self, epoch, duration):
        with self._lock:
            self._collectors[epoch].reduce_done(duration)

    def consume_batch(self, epoch, num_rows):
        with self._lock:
            self._collectors[epoch].consume_batch(num_rows)

    def consume_done(self, epoch):
        with self._lock:
            self._collectors[epoch].consume_done()

    def exchange_done(self, epoch, nbytes, seconds):
        with self._lock:
            self._collectors[epoch].exchange_done(nbytes, seconds)

    def get_stats(self, timeout: Optional[float] = None) -> TrialStats:
        if not self._trial_done_ev.wait(timeout):
            raise TimeoutError("trial did not finish")
        epoch_stats = [c.get_stats(timeout) for c in self._collectors]
        stats = TrialStats(epoch_stats, self._duration)
        return stats


# No Ray: the collector *is* the handle.
TrialStatsCollector = TrialStatsCollector_


class ObjectStoreStatsCollector:
    """Context manager sampling "store" utilisation on a background thread.
    ``bytes_used_fn`` returns bytes currently held by the shuffle (HBM arenas in
    GPU mode); default samples ``torch.cuda`` allocator + arena bytes if a CUDA
    context exists, else 0."""

    def __init__(self, utilization_sample_period=5.0,
                 bytes_used_fn: Optional[Callable[[], int]] = None,
                 do_print: bool = False):
        self._period = utilization_sample_period
        self._fn = bytes_used_fn or default_bytes_used
        self._print = do_print
        self._store_stats = None

    def __enter__(self):
        self._store_stats = []
        self._done_event = threading.Event()
        self._thread = threading.Thread(target=collect_store_stats, daemon=True,
                                        args=(self._store_stats, self._done_event,
                                              self._period, self._fn, self._print))
        self._thread.start()
        return self

    def __exit__(self, *exc_details):
        self._done_event.set()
        self._thread.join()

    def get_stats(self):
        return self._store_stats


_ARENA_BYTES_FNS: List[Callable[[], int]] = []


def register_bytes_used_source(fn: Callable[[], int]) -> None:
    """Engines register their arena byte counters here."""
    _ARENA_BYTES_FNS.append(fn)


def unregister_bytes_used_source(fn: Callable[[], int]) -> None:
    if fn in _ARENA_BYTES_FNS:
        _ARENA_BYTES_FNS.remove(fn)


_PEAK_SINCE_SAMPLE = [0]


def note_bytes_in_use(nbytes: int) -> None:
    """High-water mark between two samples: engines call this when they
    allocate, so a trial shorter than ``utilization_sample_period`` (an epoch is
    milliseconds here, the reference sampled plasma every 5 s,
    ``stats.py:258-279``) still reports what it held."""
    if nbytes > _PEAK_SINCE_SAMPLE[0]:
        _PEAK_SINCE_SAMPLE[0] = int(nbytes)


def default_bytes_used() -> int:
    total = 0
    for fn in list(_ARENA_BYTES_FNS):
        try:
            total += int(fn())
        except Exception:
            pass
    peak, _PEAK_SINCE_SAMPLE[0] = _PEAK_SINCE_SAMPLE[0], 0
    return max(total, peak)


def collect_store_stats(store_stats, done_event, utilization_sample_period,
                        bytes_used_fn=default_bytes_used, do_print=False):
    is_done = False
    while not is_done:
        get_time = timeit.default_timer()
        used = bytes_used_fn()
        if do_print:
            print(f"shuffle store in use: {human_readable_size(used)}")
        store_stats.append((get_time, StoreSample(used)))
        is_done = done_event.wait(timeout=utilization_sample_period)
    # one last sample: picks up the high-water mark of a trial that ended
    # before the first period elapsed
    store_stats.append((timeit.default_timer(), StoreSample(bytes_used_fn())))


#
# Stats processing utilities.
#

_COMMON = ["num_files", "num_row_groups_per_file", "num_reducers",
           "num_trainers", "num_epochs", "max_concurrent_epochs", "trial"]
_AGG = ("avg", "std", "max", "min")
_THROUGHPUT = ["duration", "row_throughput", "batch_throughput",
               "batch_throughput_per_trainer"]
_EXTRA = ["exchange_gbps", "exchange_roofline_frac"]
NVLINK_GBPS_PER_DIR = 900.0   # nominal NVLink 5 per direction per GPU


def _agg_names(stem: str) -> List[str]:
    return [f"{a}_{stem}" for a in _AGG]


TRIAL_FIELDS = (_COMMON + _THROUGHPUT
                + ["avg_object_store_utilization", "max_object_store_utilization"]
                + _agg_names("epoch_duration") + _agg_names("map_stage_duration")
                + _agg_names("reduce_stage_duration")
                + _agg_names("consume_stage_duration")
                + _agg_names("map_task_duration") + _agg_names("read_duration")
                + _agg_names("reduce_task_duration")
                + _agg_names("time_to_consume"))
EPOCH_FIELDS = (_COMMON + ["epoch"] + _THROUGHPUT
                + ["map_stage_duration", "reduce_stage_duration",
                   "consume_stage_duration"]
                + _agg_names("map_task_duration") + _agg_names("read_duration")
                + _agg_names("reduce_task_duration")
                + _agg_names("time_to_consume"))
CONSUMER_FIELDS = _COMMON + ["epoch", "timestamp", "num_rows_in_reducer_batch"]


def _agg(row: dict, stem: str, values) -> None:
    vals = np.asarray(list(values), dtype=np.float64)
    if vals.size == 0:
        vals = np.zeros(1)
    row[f"avg_{stem}"] = np.mean(vals)
    row[f"std_{stem}"] = np.std(vals)
    row[f"max_{stem}"] = np.max(vals)
    row[f"min_{stem}"] = np.min(vals)


def _open(filename: str, mode: str):
    if fsspec is not None:
        return fsspec.open(filename, mode=mode)
    return open(filename, mode)


def _csv_target(stats_dir, kind, hr_rows, hr_batch, unique, now, overwrite):
    name = f"{kind}_stats_{hr_rows}_rows_{hr_batch}_batch_size"
    name += f"_{now}.csv" if unique else ".csv"
    path = os.path.join(stats_dir, name)
    header = (overwrite or not os.path.exists(path) or os.path.getsize(path) == 0)
    return path, header


def _exchange_extras(epoch_stats: List[EpochStats], num_trainers: int) -> Dict[str, float]:
    nbytes = sum(e.exchange_bytes for e in epoch_stats)
    secs = sum(e.exchange_seconds for e in epoch_stats)
    if nbytes <= 0 or secs <= 0:
        return {}
    gbps = nbytes / secs / 1e9
    return {"exchange_gbps": gbps,
            "exchange_roofline_frac": gbps / (NVLINK_GBPS_PER_DIR * max(1, num_trainers))}


def process_stats(all_stats, overwrite_stats, stats_dir, no_epoch_stats,
                  no_consumer_stats, unique_stats, num_rows, num_files,
                  num_row_groups_per_file, batch_size, num_reducers,
                  num_trainers, num_epochs, max_concurrent_epochs):
    """Print the trial summary and write the trial / epoch / consumer CSVs
    (same file names, columns and append/overwrite rules as the reference)."""
    stats_list, store_stats_list = zip(*all_stats)
    times = [stats.duration for stats in stats_list]
    mean, std = np.mean(times), np.std(times)
    used = [getattr(s, "object_store_bytes_used", 0)
            for trial in store_stats_list for _, s in (trial or [])] or [0]
    nsamples = sum(len(trial or []) for trial in store_stats_list)
    rows_total = num_epochs * num_rows
    print(f"\nMean over {len(times)} trials: {mean:.3f}s +- {std}")
    print(f"Mean throughput over {len(times)} trials: "
          f"{rows_total / mean:.2f} rows/s +- "
          f"{np.std([rows_total / t for t in times]):.2f}")
    print(f"Mean batch throughput over {len(times)} trials: "
          f"{(rows_total / batch_size) / mean:.2f} batches/s +- "
          f"{np.std([(rows_total / batch_size) / t for t in times]):.2f}")
    print(f"Max object store utilization over {nsamples} "
          f"samples: {human_readable_size(np.max(used))}\n")

    if stats_dir.startswith("s3"):
        write_mode = "w"
    else:
        os.makedirs(stats_dir, exist_ok=True)
        write_mode = "w+" if overwrite_stats else "a+"
    hr_rows = human_readable_big_num(num_rows)
    hr_batch = human_readable_big_num(batch_size)
    now = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat()
    base = {"num_files": num_files,
            "num_row_groups_per_file": num_row_groups_per_file,
            "num_reducers": num_reducers, "num_trainers": num_trainers,
            "num_epochs": num_epochs,
            "max_concurrent_epochs": max_concurrent_epochs}

    def _throughput(row, rows, duration):
        row["duration"] = duration
        row["row_throughput"] = rows / duration
        row["batch_throughput"] = row["row_throughput"] / batch_size
        row["batch_throughput_per_trainer"] = row["batch_throughput"] / num_trainers

    has_extra = any(_exchange_extras(s.epoch_stats, num_trainers) for s in stats_list)

    # ---- trial stats ------------------------------------------------------
    path, header = _csv_target(stats_dir, "trial", hr_rows, hr_batch,
                               unique_stats, now, overwrite_stats)
    print(f"Writing out trial stats to {path}.")
    with _open(path, write_mode) as f:
        writer = csv.DictWriter(f, fieldnames=TRIAL_FIELDS + (_EXTRA if has_extra else []))
        if header:
            writer.writeheader()
        for trial, (stats, store_stats) in enumerate(all_stats):
            row = dict(base, trial=trial)
            _throughput(row, rows_total, stats.duration)
            ep = stats.epoch_stats
            tused = [getattr(s, "object_store_bytes_used", 0)
                     for _, s in (store_stats or [])] or [0]
            row["avg_object_store_utilization"] = np.mean(tused)
            row["max_object_store_utilization"] = np.max(tused)
            _agg(row, "epoch_duration", (e.duration for e in ep))
            _agg(row, "map_stage_duration", (e.map_stats.stage_duration for e in ep))
            _agg(row, "reduce_stage_duration",
                 (e.reduce_stats.stage_duration for e in ep))
            _agg(row, "consume_stage_duration",
                 (e.consume_stats.stage_duration for e in ep))
            _agg(row, "map_task_duration",
                 (d for e in ep for d in e.map_stats.task_durations))
            _agg(row, "read_duration",
                 (d for e in ep for d in e.map_stats.read_durations))
            _agg(row, "reduce_task_duration",
                 (d for e in ep for d in e.reduce_stats.task_durations))
            _agg(row, "time_to_consume",
                 (d for e in ep for d in e.consume_stats.time_to_consumes))
            if has_extra:
                row.update(_exchange_extras(ep, num_trainers))
            writer.writerow(row)

    # ---- epoch stats ------------------------------------------------------
    if not no_epoch_stats:
        path, header = _csv_target(stats_dir, "epoch", hr_rows, hr_batch,
                                   unique_stats, now, overwrite_stats)
        print(f"Writing out epoch stats to {path}.")
        with _open(path, write_mode) as f:
            writer = csv.DictWriter(f, fieldnames=EPOCH_FIELDS + (_EXTRA if has_extra else []))
            if header:
                writer.writeheader()
            for trial, (trial_stats, _) in enumerate(all_stats):
                for epoch, e in enumerate(trial_stats.epoch_stats):
                    row = dict(base, trial=trial, epoch=epoch)
                    _throughput(row, num_rows, e.duration)
                    row["map_stage_duration"] = e.map_stats.stage_duration
                    row["reduce_stage_duration"] = e.reduce_stats.stage_duration
                    row["consume_stage_duration"] = e.consume_stats.stage_duration
                    _agg(row, "map_task_duration", e.map_stats.task_durations)
                    _agg(row, "read_duration", e.map_stats.read_durations)
                    _agg(row, "reduce_task_duration", e.reduce_stats.task_durations)
                    _agg(row, "time_to_consume", e.consume_stats.time_to_consumes)
                    if has_extra:
                        row.update(_exchange_extras([e], num_trainers))
                    writer.writerow(row)

    # ---- consumer stats ---------------------------------------------------
    if not no_consumer_stats:
        path, header = _csv_target(stats_dir, "consumer", hr_rows, hr_batch,
                                   unique_stats, now, overwrite_stats)
        print(f"Writing out consumer stats to {path}.")
        with _open(path, write_mode) as f:
            writer = csv.DictWriter(f, fieldnames=CONSUMER_FIELDS)
            if header:
                writer.writeheader()
            for trial, (trial_stats, _) in enumerate(all_stats):
                for epoch, e in enumerate(trial_stats.epoch_stats):
                    for ts, nrows in e.consume_stats.consume_times.items():
                        writer.writerow(dict(base, trial=trial, epoch=epoch,
                                             timestamp=ts,
                                             num_rows_in_reducer_batch=nrows))


UNITS = ["", "K", "M", "B", "T", "Q"]


def human_readable_big_num(num):
    idx = int(math.log10(num) // 3) if num >= 1 else 0
    idx = min(idx, len(UNITS) - 1)
    new_num = num / 10**(3 * idx)
    if new_num % 1 == 0:
        return f"{int(new_num)}{UNITS[idx]}"
    return f"{new_num:.1f}{UNITS[idx]}"


def human_readable_size(num, precision=1, suffix="B"):
    for unit in ["", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "Zi"]:
        if abs(num) < 1024.0 or unit == "Zi":
            break
        num /= 1024.0
    return f"{num:.{precision}f}{unit}{suffix}"
