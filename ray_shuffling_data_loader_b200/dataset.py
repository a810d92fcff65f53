"""ShufflingDataset: the per-epoch shuffling iterator (components C8, C9).

Constructor signature, ``set_epoch`` contract (mandatory, must change every
epoch, ``ValueError`` otherwise), "shuffling starts at construction for the
first ``max_concurrent_epochs`` epochs", exact-``batch_size`` re-batching with
an optional short tail (``drop_last``) and the ``task_done`` accounting are the
reference's (``ray_shuffling_data_loader/dataset.py:15-205``).

Architecture differences:

* ranks are symmetric. In distributed mode (``torch.distributed`` initialised,
  world size == ``num_trainers``) every process is mapper + reducer + trainer:
  it runs its own shuffle-driver thread (the analogue of the reference's
  ``ray.remote(shuffle)`` task, ``dataset.py:68-74``) and its own queue; there
  is no rank-0 master to outlive the others (reference example sleeps 10 s for
  that, ``ray_torch_shuffle.py:248-253``). In single-process mode rank 0 hosts
  the engine for all trainers and other ranks connect to the named queue.
* reducer outputs are row ranges of one contiguous epoch buffer, so re-batching
  is pointer arithmetic (zero-copy views), and the reference's row-dropping
  re-batcher bug (``dataset.py:160-168``) cannot occur: every row is
  delivered exactly once per epoch.
* in GPU mode batches are born in HBM; the iterator yields ``DeviceBatch``
  (or pandas on the CPU backend, like the reference).
* ``state_dict``/``load_state_dict`` give mid-epoch resume (the reference has no
  checkpointing): the permutation is a pure function of
  ``(seed, epoch)``.
"""
from __future__ import annotations

import os
import threading
from typing import List, Optional

from ray_shuffling_data_loader_b200.batch_queue import BatchQueue
from ray_shuffling_data_loader_b200.parallel import bootstrap
from ray_shuffling_data_loader_b200.runtime.chunks import (DeviceBatch,
                                                           ShuffledChunk,
                                                           packed_to_dataframe,
                                                           _to_numpy)
from ray_shuffling_data_loader_b200.shuffle import BatchConsumer, shuffle

BATCHQUEUE_ACTOR_NAME = "BatchQueue"
REDUCER_CLUSTER_CORE_SHARE = 0.6


def get_num_cpus() -> int:
    return os.cpu_count() or 1


class ShufflingDataset:
    """Per-epoch globally shuffled batches of rows.

    Positional arguments (same order as reference ``dataset.py:37-45``):

    ``filenames``              input Parquet files (their concatenation is the table)
    ``num_epochs``             epochs that will be iterated
    ``num_trainers``           data-parallel consumers; each gets a disjoint 1/T of every epoch
    ``batch_size``             rows per yielded batch
    ``rank``                   which trainer this process is
    ``drop_last``              skip the final short batch (default: yield it)
    ``num_reducers``           destination chunks per epoch, each with its own completion
                               flag (default ``trainers x cores x 0.6`` like the reference)
    ``max_concurrent_epochs``  epoch ring depth (default 2)

    Shuffling of the first ``max_concurrent_epochs`` epochs starts in the
    constructor.

    Keyword-only:

    ``seed``         permutation seed (``None``: random, agreed across ranks); the same
                     seed gives the same batches on the CPU and the GPU backend
    ``backend``      ``"cuda"``, ``"cpu"`` or ``None`` (auto)
    ``output``       ``"pandas"`` (CPU default), ``"device"`` (GPU default: ``DeviceBatch``),
                     ``"packed"`` (raw ``uint8[B, pitch]``)
    ``layout_fn``    ``schema -> RowLayout`` (column projection and casts)
    ``start_epoch``  first epoch to shuffle (checkpoint resume)
    """

    def __init__(self,
                 filenames: List[str],
                 num_epochs: int,
                 num_trainers: int,
                 batch_size: int,
                 rank: int,
                 drop_last: bool = False,
                 num_reducers: int = None,
                 max_concurrent_epochs: int = 2,
                 *,
                 seed: Optional[int] = None,
                 backend: Optional[str] = None,
                 output: Optional[str] = None,
                 layout_fn=None,
                 queue_name: Optional[str] = None,
                 start_epoch: int = 0,
                 stats_collector=None,
                 chunk_wait_timeout_s: Optional[float] = None,
                 **engine_options):
        if num_reducers is None:
            num_reducers = int(
                num_trainers * get_num_cpus() * REDUCER_CLUSTER_CORE_SHARE)
        num_reducers = max(1, num_reducers)
        max_concurrent_epochs = max(1, min(max_concurrent_epochs, max(1, num_epochs)))

        self._batch_size = batch_size
        self._num_epochs = num_epochs
        self._num_trainers = num_trainers
        self._rank = rank
        self._epoch = None
        # Used to check that the user is correctly setting the epoch at the
        # beginning of each epoch.
        self._last_epoch = None
        self._drop_last = drop_last
        self._wait_timeout = chunk_wait_timeout_s
        self._skip_batches = 0
        self._batches_consumed = 0
        self._engine = None
        self._driver: Optional[threading.Thread] = None
        self._driver_error: List[BaseException] = []
        self._shuffle_duration = None

        ctx = bootstrap.current_context()
        distributed = ctx.world > 1 and ctx.world == num_trainers
        owner = distributed or rank == 0
        name = queue_name or BATCHQUEUE_ACTOR_NAME
        if distributed:
            if rank != ctx.rank:
                raise ValueError(f"rank={rank} does not match the process group "
                                 f"rank {ctx.rank}")
            name = f"{name}:{rank}"

        if owner:
            # Owner process: build the engine, the queue and kick off shuffling.
            from ray_shuffling_data_loader_b200.runtime.engine import make_engine
            # pandas batches are copies of the rows, so the host engine may reuse
            # an epoch's buffer once this iterator has released it
            engine_options.setdefault("recycle_buffers", output in (None, "pandas"))
            self._engine = make_engine(
                filenames, num_trainers=num_trainers, num_reducers=num_reducers,
                batch_size=batch_size, drop_last=drop_last, layout_fn=layout_fn,
                seed=seed, backend=backend, stats_collector=stats_collector,
                max_concurrent_epochs=max_concurrent_epochs, **engine_options)
            self._seed = self._engine.seed
            self._batch_queue = BatchQueue(
                num_epochs, num_trainers, max_concurrent_epochs,
                name=name, connect=False,
                active_ranks=self._engine.local_trainers)
            self._consumer = BatchConsumerQueue(self._batch_queue)
            # Wait until the queue has been created.
            self._batch_queue.ready()
            # Kick off shuffle on the driver thread.
            self._driver = threading.Thread(
                target=self._drive, name=f"shuffle-driver[{rank}]", daemon=True,
                args=(filenames, num_epochs, num_reducers, num_trainers,
                      stats_collector, start_epoch))
            self._driver.start()
            device = self._engine.device
        else:
            # Worker process/instance: connect to the batch queue.
            self._batch_queue = BatchQueue(
                num_epochs, num_trainers, max_concurrent_epochs,
                name=name, connect=True)
            self._seed = seed
            device = None      # decided by the first chunk that arrives
        if output is None and device is not None:
            output = "pandas" if device == "cpu" else "device"
        if output not in (None, "pandas", "device", "packed", "span"):
            raise ValueError(f"unknown output {output!r}")
        self._output = output

    # ------------------------------------------------------------------
    def _drive(self, filenames, num_epochs, num_reducers, num_trainers,
               stats_collector, start_epoch):
        try:
            self._shuffle_duration = shuffle(
                filenames, self._consumer, num_epochs, num_reducers,
                num_trainers, stats_collector=stats_collector,
                engine=self._engine, start_epoch=start_epoch)
        except BaseException as e:  # re-raised in the consumer thread
            self._driver_error.append(e)
            try:
                self._batch_queue.shutdown()
            except Exception:
                pass

    @property
    def engine(self):
        return self._engine

    @property
    def seed(self):
        return self._seed

    def set_epoch(self, epoch):
        """Select the epoch the next ``iter()`` will read; must be called with a
        new value before every pass (reference ``dataset.py:96-106``)."""
        self._epoch = epoch

    # -- checkpoint / resume (not in the reference) -------------------------
    def state_dict(self) -> dict:
        return {"seed": self._seed, "epoch": self._epoch,
                "batches_consumed": self._batches_consumed,
                "batch_size": self._batch_size}

    def load_state_dict(self, state: dict) -> None:
        """Resume inside ``state['epoch']``: construct the dataset with the same
        ``seed`` and ``start_epoch=state['epoch']``, load the state, call
        ``set_epoch(state['epoch'])`` and iterate - the first
        ``batches_consumed`` batches are skipped."""
        if self._seed is not None and state.get("seed") not in (None, self._seed):
            raise ValueError("state_dict seed does not match this dataset's seed")
        if state.get("batch_size", self._batch_size) != self._batch_size:
            raise ValueError("state_dict batch_size mismatch")
        self._skip_batches = int(state.get("batches_consumed", 0))

    def _convert(self, packed, layout):
        if self._output == "span":
            return packed           # BatchSpan, or a materialised straddling batch
        if isinstance(packed, BatchSpan):
            packed = packed.packed()
        if self._output is None:
            import numpy as np
            self._output = "pandas" if isinstance(packed, np.ndarray) else "device"
        if self._output == "packed":
            return packed
        if self._output == "device":
            return DeviceBatch(packed, layout)
        return packed_to_dataframe(_to_numpy(packed), layout)

    def _frames_apply(self, chunk: ShuffledChunk) -> bool:
        """Per-chunk DataFrames: pandas output from host (numpy) epoch buffers, from
        chunks shipped by the owning process, and - when ``output="pandas"`` is asked for
        explicitly - from device buffers (one D2H copy per chunk instead of one per batch)."""
        if self._output not in (None, "pandas"):
            return False
        import numpy as np
        data = getattr(chunk.buffer, "data", None)
        if self._output == "pandas":
            return data is not None
        if not isinstance(data, np.ndarray):
            return False
        self._output = "pandas"
        return True

    def _raise_driver_error(self):
        if self._driver_error:
            raise RuntimeError("shuffle driver failed") from self._driver_error[0]

    def __iter__(self):
        """
        This iterator yields batches from the shuffling queue.
        """
        if self._epoch is None or self._epoch == self._last_epoch:
            raise ValueError(
                "You must set the epoch on this dataset via set_epoch()"
                "at the beginning of each epoch, before iterating over this "
                "dataset (e.g. via enumerate(ds)).")
        epoch = self._epoch
        rebatch = _Rebatcher(self._batch_size)
        frames = _ChunkFrames()
        self._batches_consumed = 0
        skip = self._skip_batches
        self._skip_batches = 0
        buffers = {}
        layout = None
        is_done = False
        finished = False
        unacked = 0
        try:
            while not is_done:
                # Get a batch of reducer chunks from the queue.
                try:
                    pending = self._batch_queue.get_batch(self._rank, epoch)
                except Exception:
                    self._raise_driver_error()
                    raise
                if pending and pending[-1] is None:
                    # Set done flag but don't break yet, since we might still
                    # have more items in pending to consume.
                    is_done = True
                    pending.pop()
                num_outstanding = unacked = len(pending)
                for chunk in pending:
                    chunk.wait(self._wait_timeout)
                    layout = chunk.layout
                    buffers[id(chunk.buffer)] = chunk.buffer
                    rebatch.push(chunk)
                    by_frames = self._frames_apply(chunk)
                    if by_frames:
                        frames.add(chunk)
                    for packed in rebatch.pop_full():
                        self._batches_consumed += 1
                        if by_frames and isinstance(packed, BatchSpan):
                            df = frames.take(packed)      # also when skipped: prunes frames
                            if skip > 0:
                                skip -= 1
                                continue
                            yield df
                            continue
                        if skip > 0:
                            skip -= 1
                            continue
                        yield self._convert(packed, layout)
                if num_outstanding > 0:
                    # Signal to the queue that we're done with these chunks.
                    self._batch_queue.task_done(self._rank, epoch, num_outstanding)
                    unacked = 0
            # Yield leftover (incomplete) batch if we're not dropping
            # incomplete batches.
            tail = rebatch.pop_tail()
            if tail is not None and not self._drop_last:
                self._batches_consumed += 1
                if skip <= 0:
                    if frames.frames and isinstance(tail, BatchSpan):
                        yield frames.take(tail)
                    else:
                        yield self._convert(tail, layout)
            finished = True
        finally:
            if unacked:
                # Abandoned mid-group (early ``break``): acknowledge what we hold.
                try:
                    self._batch_queue.task_done(self._rank, epoch, unacked)
                except Exception:
                    pass
            if not finished and not is_done:
                self._drain(epoch)
            for buf in buffers.values():
                buf.release()
            if finished or is_done:
                # Account for the producer_done sentinel.
                try:
                    self._batch_queue.task_done(self._rank, epoch, 1)
                except Exception:
                    pass
            self._last_epoch = epoch
        if epoch == self._num_epochs - 1:
            self._finish()

    def _drain(self, epoch):
        """The consumer abandoned the epoch early: swallow the rest so the
        epoch window (and every other rank) keeps moving."""
        try:
            while True:
                pending = self._batch_queue.get_batch(self._rank, epoch)
                done = bool(pending) and pending[-1] is None
                self._batch_queue.task_done(self._rank, epoch, len(pending))
                if done:
                    return
        except Exception:
            return

    def _finish(self):
        if self._engine is None:
            return      # a connected worker never owns the queue
        if self._driver is not None:
            # Returns once every trainer has consumed the final epoch.
            self._driver.join()
            self._raise_driver_error()
            self._driver = None
        # Batches the caller still holds stay valid: device memory is released
        # by close() / garbage collection, not here.
        self._engine.quiesce()
        try:
            self._batch_queue.shutdown()
        except Exception:
            pass

    def close(self):
        """Tear down early (the last epoch's iterator does this itself)."""
        if self._engine is None:
            return
        if hasattr(self._engine, "cancel") and self._driver is not None:
            self._engine.cancel()
        try:
            self._batch_queue.shutdown()
        except Exception:
            pass
        if self._driver is not None:
            self._driver.join(timeout=60)
            self._driver = None
        if self._engine is not None:
            self._engine.close()

    def __del__(self):
        try:
            if self._engine is not None and self._driver is None:
                self._engine.close()
        except Exception:
            pass


class _ChunkFrames:
    """pandas output on the host backend: every reducer chunk is turned into a
    DataFrame ONCE, as soon as it is complete (one multi-threaded transposition of the
    packed rows, ``runtime.chunks.unpack_columns``), and batches are row slices of those
    frames - views inside a chunk, a two-piece ``concat`` across a chunk boundary. This
    is what the reference does with its reducer outputs (``dataset.py:144-168``);
    converting packed rows batch by batch cost 80 % of the CPU path's time."""

    def __init__(self):
        self.frames = []            # [buffer, row_start, row_stop, DataFrame], in order

    def add(self, chunk: ShuffledChunk):
        if len(chunk) == 0:
            return
        packed = chunk.buffer.view(chunk.row_start, chunk.row_stop)
        self.frames.append([chunk.buffer, chunk.row_start, chunk.row_stop,
                            packed_to_dataframe(_to_numpy(packed), chunk.layout)])

    def take(self, span: "BatchSpan"):
        import pandas as pd
        pieces = []
        for buf, a, b, df in self.frames:
            if buf is span.buffer and a < span.stop and b > span.start:
                lo, hi = max(a, span.start), min(b, span.stop)
                pieces.append(df.iloc[lo - a:hi - a])
        # frames that lie entirely before this batch will never be needed again
        self.frames = [f for f in self.frames
                       if not (f[0] is span.buffer and f[2] <= span.stop)]
        if len(pieces) == 1:
            return pieces[0].reset_index(drop=True)
        return pd.concat(pieces, ignore_index=True)


class BatchSpan:
    """Rows ``[start, stop)`` of one epoch buffer: a batch before any view is
    built. ``output="span"`` yields these so wrappers can slice cached typed
    views of the whole buffer instead of re-deriving them per batch."""
    __slots__ = ("buffer", "start", "stop")

    def __init__(self, buffer, start, stop):
        self.buffer, self.start, self.stop = buffer, start, stop

    def __len__(self):
        return self.stop - self.start

    def packed(self):
        return self.buffer.view(self.start, self.stop)


class _Rebatcher:
    """Carve exact ``batch_size`` batches out of successive chunks.

    Adjacent chunks of the same epoch buffer are merged into one span, so a
    batch that straddles a chunk boundary is still a zero-copy view; only chunks
    living in different buffers (a remote, pickled chunk) are concatenated."""

    def __init__(self, batch_size: int):
        self.batch_size = batch_size
        self.spans = []     # [buffer, start, stop]
        self.rows = 0

    def push(self, chunk: ShuffledChunk):
        if len(chunk) == 0:
            return
        if self.spans and self.spans[-1][0] is chunk.buffer \
                and self.spans[-1][2] == chunk.row_start:
            self.spans[-1][2] = chunk.row_stop
        else:
            self.spans.append([chunk.buffer, chunk.row_start, chunk.row_stop])
        self.rows += len(chunk)

    def _take(self, n: int):
        """-> ``BatchSpan`` (rows of one epoch buffer, the zero-copy common case)
        or a materialised array when the batch straddles different buffers."""
        pieces = []
        while n > 0:
            buf, start, stop = self.spans[0]
            k = min(n, stop - start)
            pieces.append(BatchSpan(buf, start, start + k))
            if start + k == stop:
                self.spans.pop(0)
            else:
                self.spans[0][1] = start + k
            n -= k
            self.rows -= k
        if len(pieces) == 1:
            return pieces[0]
        mats = [p.packed() for p in pieces]
        import numpy as np
        if isinstance(mats[0], np.ndarray):
            return np.concatenate(mats, axis=0)
        import torch
        return torch.cat(mats, dim=0)

    def pop_full(self):
        while self.rows >= self.batch_size:
            yield self._take(self.batch_size)

    def pop_tail(self):
        return self._take(self.rows) if self.rows > 0 else None


class BatchConsumerQueue(BatchConsumer):
    def __init__(self, batch_queue: BatchQueue):
        self._batch_queue = batch_queue

    def consume(self, rank: int, epoch: int, batches: List[ShuffledChunk]):
        self._batch_queue.put_batch(rank, epoch, batches)

    def producer_done(self, rank: int, epoch: int):
        self._batch_queue.producer_done(rank, epoch)

    def wait_until_ready(self, epoch: int):
        self._batch_queue.new_epoch(epoch)

    def wait_until_all_epochs_done(self):
        self._batch_queue.wait_until_all_epochs_done()


def _smoke_main(argv=None) -> int:
    """``python -m ray_shuffling_data_loader_b200.dataset``: generate a small
    ``DATA_SPEC`` table, iterate it for a few epochs and *verify* delivery (role of
    the reference's ``__main__`` block, ``dataset.py:208-252``, which only prints):
    every key exactly once per epoch, exact batch sizes, a new order per epoch."""
    import argparse
    import tempfile
    import numpy as np
    from ray_shuffling_data_loader_b200.data_generation import generate_data
    ap = argparse.ArgumentParser(description=_smoke_main.__doc__)
    ap.add_argument("--num-rows", type=int, default=10**6)
    ap.add_argument("--num-files", type=int, default=10)
    ap.add_argument("--num-epochs", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=20000)
    ap.add_argument("--num-reducers", type=int, default=8)
    ap.add_argument("--backend", default=None, choices=[None, "cpu", "cuda"])
    a = ap.parse_args(argv)
    with tempfile.TemporaryDirectory() as data_dir:
        files, nbytes = generate_data(a.num_rows, a.num_files, 1, 0.0, data_dir)
        print(f"{len(files)} files, {a.num_rows} rows ({nbytes / 1e6:.1f} MB decoded); "
              f"{a.num_epochs} epochs, batch {a.batch_size}, {a.num_reducers} reducers")
        ds = ShufflingDataset(files, a.num_epochs, 1, a.batch_size, 0,
                              num_reducers=a.num_reducers, backend=a.backend, output="pandas")
        orders = []
        for epoch in range(a.num_epochs):
            ds.set_epoch(epoch)
            keys = []
            for batch in ds:
                if len(batch) != a.batch_size and len(keys) * a.batch_size + len(batch) != a.num_rows:
                    raise AssertionError("short batch in the middle of an epoch")
                keys.append(batch["key"].to_numpy())
            keys = np.concatenate(keys)
            if not np.array_equal(np.sort(keys), np.arange(a.num_rows)):
                raise AssertionError(f"epoch {epoch}: rows lost or duplicated")
            orders.append(keys[:64].copy())
            print(f"epoch {epoch}: {len(keys)} rows, exactly once")
        if len(orders) > 1 and all(np.array_equal(orders[0], o) for o in orders[1:]):
            raise AssertionError("epochs were not reshuffled")
    print("ok")
    return 0


if __name__ == "__main__":
    raise SystemExit(_smoke_main())
