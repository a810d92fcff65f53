"""NCCL ``all_to_all_single`` row exchange - the baseline, not the product.

BASELINE.json: "A path that only calls NCCL all-to-all for the row exchange is
the baseline". This is that path, written the way one would write it with
library collectives and no fused kernel - and written *well*, so the comparison
with the fused P2P scatter is fair:

    1. pack    the scatter kernel with an identity key writes this rank's rows,
               cast + packed, into a staging buffer in source order
    2. route   ``perm_positions`` -> (trainer, slot) per row; one radix sort by
               trainer groups rows per destination
    3. gather  one permuted row copy builds fixed-size send blocks
               ``[world][cap]`` (rows + their destination offsets; unused tail
               entries carry offset -1)
    4. NCCL    two equal-split ``all_to_all_single`` (rows, offsets)
    5. place   one permuted row copy drops received rows at their final slots

Nothing in an epoch synchronises with the host: block capacities are fixed
(``mean + 12 sigma`` of the per-destination row count, so an overflow is a
< 1e-30 event and the bench's exactly-once check would expose it), every
staging buffer is allocated once, and the whole sequence is enqueued on a
dedicated low-priority stream that the engine's shuffle stream joins. (Round 1's
version called ``.tolist()`` twice and ``stream.synchronize()`` once per epoch
and exchanged variable-size blocks; that sandbagged the baseline.)

Compared with the product path this still costs three extra HBM round trips of
the whole epoch and two collectives (reference analogue: the implicit mapper ->
reducer all-to-all through the object store, ``shuffle.py:120-123``).
"""
from __future__ import annotations

import math
from typing import List


class _State:
    """Per-engine staging buffers and the exchange stream (allocated once)."""

    def __init__(self, engine):
        import torch
        dev = torch.device("cuda", engine.device_index)
        n, pitch, world = max(engine.n_local, 1), engine.layout.row_pitch, engine.world
        self.stream = torch.cuda.Stream(device=dev, priority=0)
        self.packed = torch.empty((n, pitch), dtype=torch.uint8, device=dev)
        self.trainer = torch.empty(n, dtype=torch.int32, device=dev)
        self.slots = torch.empty(n, dtype=torch.int64, device=dev)
        self.arange = torch.arange(n, dtype=torch.int64, device=dev)
        if world > 1:
            # block capacity must be IDENTICAL on every rank (equal-split collective):
            # derive it from the largest source range, not from this rank's row count
            n_max = -(-engine.plan.num_rows // world)
            mean = n_max / world
            self.cap = int(math.ceil(mean + 12.0 * math.sqrt(mean) + 1024))
            blk = world * self.cap
            self.send_rows = torch.empty((blk, pitch), dtype=torch.uint8, device=dev)
            self.recv_rows = torch.empty((blk, pitch), dtype=torch.uint8, device=dev)
            self.send_off = torch.empty(blk + 1, dtype=torch.int64, device=dev)
            self.recv_off = torch.empty(blk, dtype=torch.int64, device=dev)


def _state(engine) -> _State:
    st = getattr(engine, "_nccl_state", None)
    if st is None:
        st = engine._nccl_state = _State(engine)
    return st


def exchange_epoch(engine, key_words: List[int], slot: int) -> None:
    """Fill this rank's epoch slot(s) for the epoch described by ``key_words``.
    Enqueued on the baseline's own stream; the engine's shuffle stream joins it."""
    import torch
    import torch.distributed as dist
    C, plan, lay = engine.C, engine.plan, engine.layout
    n, pitch = engine.n_local, lay.row_pitch
    if engine.resident != "hbm":
        raise NotImplementedError("the NCCL baseline needs resident='hbm'")
    st = _state(engine)
    stream = st.stream
    s = stream.cuda_stream
    # order after whatever the shuffle stream did before (slot reuse gate)
    ev_in = C.event_create(False)
    C.event_record(ev_in, engine.shuffle_stream)
    C.stream_wait_event(s, ev_in)
    C.event_destroy(ev_in)

    with torch.cuda.stream(stream):
        if n:
            # 1. pack in source order: identity permutation, single "trainer"
            saved = engine.shuffle_stream
            engine.shuffle_stream = s
            try:
                _launch_identity(engine, [1, 1, 1, 0, 0, 0, 0, 0, 0], st.packed.data_ptr(), n)
            finally:
                engine.shuffle_stream = saved
            # 2. route
            C.perm_positions(key_words, plan.num_rows, plan.num_trainers, engine.src_lo, n,
                             st.trainer.data_ptr(), st.slots.data_ptr(), s)
            engine.launches += 1
        trainer, slots = st.trainer[:n], st.slots[:n]
        if engine.world == 1:
            # every trainer is local: destination offset inside the arena's slot block
            if n:
                nloc = plan.num_trainers
                base = engine._slot_ptr(slot, 0)
                off = trainer.to(torch.int64) * engine.slot_bytes + slots * pitch
                C.place_rows(st.packed.data_ptr(), off.data_ptr(), n, pitch, base, s)
                engine.launches += 1
                assert engine._slot_ptr(slot, nloc - 1) == base + (nloc - 1) * engine.slot_bytes
        else:
            world, cap = engine.world, st.cap
            blk = world * cap
            sorted_t, order = torch.sort(trainer)                       # radix sort, 1 pass
            counts = torch.bincount(trainer, minlength=world)
            starts = torch.cumsum(counts, 0) - counts
            st64 = sorted_t.to(torch.int64)
            in_block = st.arange[:n] - starts[st64]
            fits = in_block < cap                 # (false only beyond mean + 12 sigma)
            send_index = st64 * cap + in_block
            # 3. gather into fixed-size blocks: rows and their destination byte offsets;
            #    entry ``blk`` of send_off is a trash can for rows that do not fit
            st.send_off.fill_(-1)
            st.send_off.index_copy_(0, torch.where(fits, send_index, blk),
                                    slots.index_select(0, order) * pitch)
            row_off = torch.where(fits, send_index * pitch, -1)
            C.place_rows(st.packed.data_ptr(), row_off.data_ptr(), n, pitch,
                         st.send_rows.data_ptr(), s, src_idx=order.data_ptr())
            engine.launches += 1
            # 4. the collectives (equal splits: no sizes to agree on, no host sync)
            dist.all_to_all_single(st.recv_rows, st.send_rows, group=engine.exchange_pg)
            dist.all_to_all_single(st.recv_off, st.send_off[:blk], group=engine.exchange_pg)
            # 5. place
            C.place_rows(st.recv_rows.data_ptr(), st.recv_off.data_ptr(), blk, pitch,
                         engine._slot_ptr(slot, engine.rank), s)
            engine.launches += 1
            # (temporaries were allocated with ``stream`` current: the caching
            # allocator already ties their reuse to it)
    # hand back to the shuffle stream (it publishes the produced flag next)
    ev_out = C.event_create(False)
    C.event_record(ev_out, s)
    C.stream_wait_event(engine.shuffle_stream, ev_out)
    C.event_destroy(ev_out)


def _launch_identity(engine, ident_key, dst_ptr: int, n: int) -> None:
    """Pack rows [0, n) of the resident table to ``dst_ptr`` in source order."""
    plan_rows = max(n, 1)
    saved_plan = engine.plan
    from ray_shuffling_data_loader_b200.ops.plan import ShufflePlan
    engine.plan = ShufflePlan(plan_rows, 1, 1, saved_plan.batch_size)
    try:
        engine._launch_chunk(ident_key, 0, n, 0, [dst_ptr])
    finally:
        engine.plan = saved_plan
