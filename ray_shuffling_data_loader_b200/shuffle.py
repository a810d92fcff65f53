"""Shuffle driver and epoch builder (components C1, C2, C5).

Public contract kept from the reference (``ray_shuffling_data_loader/shuffle.py``):

* ``BatchConsumer`` - the 4-method sink protocol (``shuffle.py:11-43``);
* ``shuffle(filenames, batch_consumer, num_epochs, num_reducers, num_trainers,
  stats_collector=None) -> duration`` (``shuffle.py:51-86``): per epoch, block on
  the consumer's back-pressure, launch the epoch (non-blocking), finally wait
  until everything was consumed and return the wall-clock seconds;
* ``shuffle_epoch`` hands each trainer its list of reducer outputs followed by
  ``producer_done`` (``shuffle.py:89-126,203-219``).

What an "epoch launch" is differs completely. The reference builds a Ray task
DAG - one ``shuffle_map`` per file re-reading Parquet and making R mask passes,
one ``shuffle_reduce`` per reducer doing concat + ``sample(frac=1)`` - whose
outputs cross the object store twice. Here ``engine.start_epoch`` enqueues one
fused scatter kernel per epoch on a side stream (``csrc/shuffle_kernels.cu``):
every source row is pushed, already cast and packed, straight into its final
``(trainer, slot)`` over NVLink. "Reducer outputs" are row ranges of the
destination epoch buffer (``ShuffledChunk``); nothing is materialised twice.
"""
from __future__ import annotations

import timeit
from typing import List, Optional, Sequence, Union

from ray_shuffling_data_loader_b200.runtime.chunks import ShuffledChunk
from ray_shuffling_data_loader_b200.stats import TrialStatsCollector


class BatchConsumer:
    """Sink protocol of ``shuffle()`` - the four hooks the reference defines
    (``shuffle.py:11-43``); method names and argument order are the contract, so a
    consumer written for the reference plugs in unchanged.

    * ``consume(rank, epoch, batches)``      trainer ``rank`` receives the epoch's reducer
                                             chunks (``ShuffledChunk`` handles, possibly not
                                             yet complete: call ``.wait()`` before reading)
    * ``producer_done(rank, epoch)``         no more chunks will come for that trainer/epoch
    * ``wait_until_ready(epoch)``            back-pressure: return when the shuffle of
                                             ``epoch`` may start
    * ``wait_until_all_epochs_done()``       return when everything handed out was consumed
    """

    def consume(self, rank, epoch, batches):
        raise NotImplementedError(f"{type(self).__name__} must implement consume()")

    def producer_done(self, rank, epoch):
        raise NotImplementedError(f"{type(self).__name__} must implement producer_done()")

    def wait_until_ready(self, epoch):
        raise NotImplementedError(f"{type(self).__name__} must implement wait_until_ready()")

    def wait_until_all_epochs_done(self):
        raise NotImplementedError(
            f"{type(self).__name__} must implement wait_until_all_epochs_done()")


#
# Resident shuffling: data is decoded once, every epoch is device work only.
#


def shuffle(
        filenames: Sequence[str],
        batch_consumer: BatchConsumer,
        num_epochs: int,
        num_reducers: int,
        num_trainers: int,
        stats_collector: Union[TrialStatsCollector, None] = None,
        *,
        engine=None,
        seed: Optional[int] = None,
        backend: Optional[str] = None,
        batch_size: Optional[int] = None,
        layout_fn=None,
        start_epoch: int = 0,
        **engine_options,
) -> float:
    """
    Shuffle the provided dataset every epoch.

    Args:
        filenames (str): Paths to input Parquet files.
        batch_consumer (BatchConsumer): Consumer of shuffle outputs.
        num_epochs (int): Number of training epochs.
        num_reducers (int): The number of shuffler reducers: destination
            chunks handed to ``consume`` as separate items. On the GPU engine a
            chunk's ``wait()`` returns as soon as the destination-chunk pass that
            delivers it has completed (``chunk_passes`` engine option, K7), not
            when the whole epoch has.
        num_trainers (int): Number of trainer workers.
        stats_collector(Optional[TrialStatsCollector]): Shuffle stats
            collector.
        engine: a prebuilt engine (``runtime.engine.make_engine``); built from
            the remaining keyword arguments when omitted.
        seed: permutation seed; ``None`` draws one (agreed across ranks).
        start_epoch: first epoch to produce (checkpoint resume).

    Returns:
        Wall-clock seconds, including waiting for the last epoch to be consumed.
    """
    own_engine = engine is None
    if own_engine:
        from ray_shuffling_data_loader_b200.runtime.engine import make_engine
        engine = make_engine(
            filenames, num_trainers=num_trainers, num_reducers=num_reducers,
            batch_size=batch_size or (1 << 62), layout_fn=layout_fn, seed=seed,
            backend=backend, stats_collector=stats_collector, **engine_options)
    start = timeit.default_timer()
    try:
        for epoch_idx in range(start_epoch, num_epochs):
            # Wait until consumer is ready for another epoch shuffle to start.
            throttle_start = timeit.default_timer()
            batch_consumer.wait_until_ready(epoch_idx)
            if stats_collector is not None:
                stats_collector.epoch_throttle_done(
                    epoch_idx, timeit.default_timer() - throttle_start)

            shuffle_epoch(epoch_idx, engine, batch_consumer, stats_collector)

        batch_consumer.wait_until_all_epochs_done()
    finally:
        if own_engine:
            engine.close()
    end = timeit.default_timer()
    duration = end - start

    if stats_collector is not None:
        stats_collector.trial_done(duration)

    return duration


def shuffle_epoch(epoch: int, engine, batch_consumer: BatchConsumer,
                  stats_collector: Union[TrialStatsCollector, None] = None) -> None:
    """
    Launch the shuffle for the specified epoch and hand the (not yet
    necessarily complete) reducer chunks to the consumer.
    """
    if stats_collector is not None:
        stats_collector.epoch_start(epoch)
    buffers = engine.start_epoch(epoch)
    plan = engine.plan
    for rank in engine.local_trainers:
        buf = buffers[rank]
        chunks = [ShuffledChunk(buf, i, a, b)
                  for i, (a, b) in enumerate(plan.trainer_chunks(rank))]
        consume(rank, batch_consumer, epoch, chunks)


def consume(rank: int, batch_consumer: BatchConsumer, epoch: int,
            batches: List[ShuffledChunk]) -> None:
    """
    Consume the provided batches. This is the sink of the shuffle.
    """
    batch_consumer.consume(rank, epoch, batches)
    # Signal to batch consumer that we're done producing batches for this
    # epoch.
    batch_consumer.producer_done(rank, epoch)
