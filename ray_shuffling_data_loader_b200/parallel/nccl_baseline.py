"""NCCL ``all_to_all_single`` row exchange - the baseline, not the product.

BASELINE.json: "A path that only calls NCCL all-to-all for the row exchange is
the baseline". This is that path, built from library collectives plus the
minimum of our own kernels, so the fused P2P scatter has an honest same-hardware
comparator (and a second implementation to cross-check results against):

    1. pack   - the scatter kernel with an *identity* key writes this rank's rows,
                cast + packed, into a temporary in source order        (1 pass)
    2. route  - ``perm_positions`` gives (trainer, slot) per row; a stable sort
                by trainer groups rows per destination                  (1 pass)
    3. NCCL   - ``all_to_all_single`` for the rows and for their slots
    4. place  - ``place_rows`` scatters received rows to their final slots
                                                                        (1 pass)

i.e. three extra HBM round trips and two collectives per epoch, versus one fused
kernel on the product path (reference analogue: shuffle.py:120-123, the implicit
mapper -> reducer all-to-all through the object store).
"""
from __future__ import annotations

from typing import List


def exchange_epoch(engine, key_words: List[int], slot: int) -> None:
    """Fill this rank's epoch slot(s) for the epoch described by ``key_words``.
    Runs on torch's current stream; the engine's shuffle stream waits for it."""
    import torch
    import torch.distributed as dist
    C, plan, lay = engine.C, engine.plan, engine.layout
    dev = torch.device("cuda", engine.device_index)
    n, pitch = engine.n_local, lay.row_pitch
    stream = torch.cuda.current_stream(dev)
    s = stream.cuda_stream
    if engine.resident != "hbm":
        raise NotImplementedError("the NCCL baseline needs resident='hbm'")
    # order after whatever the shuffle stream did before (slot reuse)
    ev_in = C.event_create(False)
    C.event_record(ev_in, engine.shuffle_stream)
    C.stream_wait_event(s, ev_in)
    C.event_destroy(ev_in)

    packed = torch.empty((max(n, 1), pitch), dtype=torch.uint8, device=dev)
    trainer = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    slots = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    if n:
        # 1. pack in source order: identity permutation, single "trainer"
        ident = [1, 1, 1, 0, 0, 0, 0, 0, 0]
        saved = engine.shuffle_stream
        engine.shuffle_stream = s
        try:
            _launch_identity(engine, ident, packed.data_ptr(), n)
        finally:
            engine.shuffle_stream = saved
        # 2. route
        C.perm_positions(key_words, plan.num_rows, plan.num_trainers, engine.src_lo, n,
                         trainer.data_ptr(), slots.data_ptr(), s)
        engine.launches += 1
    packed, trainer, slots = packed[:n], trainer[:n], slots[:n]
    if engine.world == 1:
        for t in engine.local_trainers:
            sel = (trainer == t).nonzero(as_tuple=True)[0]
            rows = packed.index_select(0, sel)
            sl = slots.index_select(0, sel).contiguous()
            if rows.shape[0]:
                C.place_rows(rows.data_ptr(), sl.data_ptr(), rows.shape[0], pitch,
                             engine._slot_ptr(slot, t), s)
                engine.launches += 1
            rows.record_stream(stream)
    else:
        order = torch.argsort(trainer, stable=True)
        counts = torch.bincount(trainer, minlength=engine.world).to(torch.int64)
        send_rows = packed.index_select(0, order)
        send_slots = slots.index_select(0, order)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=engine.pg)
        in_split = counts.tolist()
        out_split = recv_counts.tolist()
        total = int(sum(out_split))
        recv_rows = torch.empty((total, pitch), dtype=torch.uint8, device=dev)
        recv_slots = torch.empty(total, dtype=torch.int64, device=dev)
        # 3. the collective(s)
        dist.all_to_all_single(recv_rows, send_rows, out_split, in_split, group=engine.pg)
        dist.all_to_all_single(recv_slots, send_slots, out_split, in_split, group=engine.pg)
        # 4. place
        if total:
            C.place_rows(recv_rows.data_ptr(), recv_slots.data_ptr(), total, pitch,
                         engine._slot_ptr(slot, engine.rank), s)
            engine.launches += 1
    # hand back to the shuffle stream (it publishes the produced flag next)
    ev_out = C.event_create(False)
    C.event_record(ev_out, s)
    C.stream_wait_event(engine.shuffle_stream, ev_out)
    C.event_destroy(ev_out)
    stream.synchronize()    # temporaries are torch-allocated: keep lifetime simple


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
