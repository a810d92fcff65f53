"""CPU (numpy) shuffle engine: the semantic reference for the device engine.

Implements exactly the pipeline the CUDA engine implements - ingest the rows
this process owns once, then per epoch evaluate ``pi_e`` for every local row,
split positions into (trainer, slot) and write the cast+packed row into the
destination trainer's epoch buffer - with numpy on the host and, when several
processes cooperate, a gloo ``all_to_all`` for the row exchange. It is both the
no-GPU backend (BASELINE.json config 1: "ShufflingDataset num_trainers=1
num_reducers=2 on CPU") and the golden model the GPU tests diff against
byte for byte.

Replaces reference ``shuffle_map``/``shuffle_reduce`` (``shuffle.py:129-200``):
no per-epoch Parquet re-read, no R boolean-mask passes, no concat, no second
permutation pass.
"""
from __future__ import annotations

import os
import threading
import timeit
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.ops import perm
from ray_shuffling_data_loader_b200.ops.plan import ShufflePlan
from ray_shuffling_data_loader_b200.runtime import ingest
from ray_shuffling_data_loader_b200.runtime.chunks import EpochBuffer


def _load_native():
    """The C++ runtime (``csrc/bindings.cpp``: worker pool + the shared
    ``perm.cuh`` bijection) when it has been built; ``None`` keeps the pure
    numpy path, which is also the golden model of the native one."""
    if os.environ.get("RSDL_CPU_NATIVE", "1") == "0":
        return None
    try:
        from ray_shuffling_data_loader_b200 import _C
        return _C if hasattr(_C, "host_scatter_rows") else None
    except ImportError:
        return None


_FIELD_DTYPE = np.dtype([("src", "<u8"), ("src_code", "<u4"), ("dst_code", "<u4"),
                         ("dst_off", "<u4"), ("width", "<u4")])


def native_pack_rows(C, pool, columns: Dict[str, np.ndarray], layout: L.RowLayout) -> np.ndarray:
    """``L.pack_rows`` (all rows, in order) on the C++ worker pool: one pass
    over row blocks instead of one strided numpy assignment per column.
    Layouts with fp8 fields keep the numpy path (block scaling lives there)."""
    if layout.scale_offset >= 0 or any(L.DT_FP8 in (f.src_code, f.dst_code) for f in layout.fields):
        return L.pack_rows(columns, layout)
    n = len(columns[layout.fields[0].name])
    out = np.empty((n, layout.row_pitch), dtype=np.uint8)
    desc = np.zeros(len(layout.fields), dtype=_FIELD_DTYPE)
    keep = []
    for i, f in enumerate(layout.fields):
        col = np.ascontiguousarray(columns[f.name], dtype=L.numpy_storage_dtype(f.src_code))
        if col.size != n * f.width:
            raise ValueError(f"column {f.name}: {col.shape} does not hold {n} x {f.width}")
        keep.append(col)
        desc[i] = (col.ctypes.data, f.src_code, f.dst_code, f.offset, f.width)
    if n:
        C.host_pack_rows(pool, desc.ctypes.data, len(desc), n, layout.row_pitch, out.ctypes.data)
    return out


class CpuShuffleEngine:
    """See module docstring. ``world``/``rank`` describe cooperating processes
    (gloo); with ``world == 1`` this process serves every trainer."""

    device = "cpu"

    def __init__(self, filenames: Sequence[str], plan_args: dict,
                 layout_fn, seed: int, rank: int = 0, world: int = 1,
                 stats_collector=None, num_threads: Optional[int] = None,
                 process_group=None, index: Optional[ingest.DatasetIndex] = None,
                 native: Optional[bool] = None, recycle_buffers: bool = False):
        self.index = index or ingest.scan_files(filenames)
        self.C = _load_native() if native in (None, True) else None
        if native is True and self.C is None:
            raise RuntimeError("native=True but ray_shuffling_data_loader_b200._C is not built")
        self.plan = ShufflePlan(num_rows=self.index.num_rows, **plan_args)
        if world > 1 and self.plan.num_trainers != world:
            raise ValueError("distributed mode needs num_trainers == world size")
        self.layout: L.RowLayout = layout_fn(self.index.schema)
        self.seed = int(seed)
        self.rank = rank
        self.world = world
        self.pg = process_group
        self.stats = stats_collector
        self.num_threads = num_threads or max(1, min(8, (os.cpu_count() or 2)))
        self.local_trainers: List[int] = ([rank] if world > 1
                                          else list(range(self.plan.num_trainers)))
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="cpu-shuffle")
        self._host_pool = self.C.HostPool(self.num_threads) if self.C is not None else None
        self._packed: Optional[np.ndarray] = None
        # Released epoch buffers are reused (the host analogue of the device
        # epoch ring) when the consumer only ever sees copies of the rows - the
        # pandas output path; a fresh multi-GB buffer per epoch costs one page
        # fault per 4 KB before the first row lands.
        self._recycle = bool(recycle_buffers)
        self._free: Dict[tuple, List[np.ndarray]] = {}
        self._lock = threading.Lock()
        self._ingest_reads: List[float] = []
        self._bytes_in_flight = 0
        self._closed = False
        from ray_shuffling_data_loader_b200 import stats as stats_mod
        self._bytes_fn = self.bytes_in_use
        stats_mod.register_bytes_used_source(self._bytes_fn)

    # -- ingest -------------------------------------------------------------
    def _ensure_ingested(self, epoch: int):
        with self._lock:
            if self._packed is not None:
                if self.stats is not None:
                    for _ in range(len(self.index.filenames)):
                        self.stats.map_start(epoch)
                        self.stats.map_done(epoch, 0.0, 0.0)
                return
            lo, hi = self.plan.source_range(self.rank, self.world)
            t0 = timeit.default_timer()
            if self.stats is not None:
                for _ in range(len(self.index.filenames)):
                    self.stats.map_start(epoch)
            table = ingest.load_table(self.index, lo, hi, columns=self.layout.names,
                                      num_threads=self.num_threads)
            # Cast + pack once; every epoch is then a pure row permutation.
            if self.C is not None:
                self._packed = native_pack_rows(self.C, self._host_pool, table.columns, self.layout)
            else:
                self._packed = L.pack_rows(table.columns, self.layout)
            self._offset = lo
            dur = timeit.default_timer() - t0
            reads = table.read_durations or [0.0]
            if self.stats is not None:
                nfiles = len(self.index.filenames)
                for i in range(nfiles):
                    self.stats.map_done(epoch, dur / nfiles,
                                        float(np.mean(reads)))

    # -- one epoch ------------------------------------------------------------
    def start_epoch(self, epoch: int) -> Dict[int, EpochBuffer]:
        """Non-blocking: returns the epoch buffers (one per local trainer);
        they become ready when the background shuffle has filled them."""
        plan, layout = self.plan, self.layout
        buffers = {}
        for t in self.local_trainers:
            rows = plan.trainer_rows(t)
            shape = (rows, layout.row_pitch)
            data = None
            if self._recycle:
                with self._lock:
                    pool = self._free.get(shape)
                    data = pool.pop() if pool else None
            if data is None:
                data = np.empty(shape, dtype=np.uint8)
            buffers[t] = EpochBuffer(epoch, t, rows, layout, data, "cpu",
                                     release_fn=(self._make_recycler(data)
                                                 if self._recycle else None))
            # (a CPU buffer with a release_fn still waits on its ready event)
            self._bytes_in_flight += data.nbytes
        from ray_shuffling_data_loader_b200 import stats as stats_mod
        stats_mod.note_bytes_in_use(self.bytes_in_use())
        self._pool.submit(self._run_epoch, epoch, buffers)
        return buffers

    def _make_recycler(self, data: np.ndarray):
        def recycle():
            if self._closed:
                return
            with self._lock:
                pool = self._free.setdefault(data.shape, [])
                if len(pool) < 4:                 # never hoard more than a few epochs
                    pool.append(data)
            self._bytes_in_flight -= data.nbytes
        return recycle

    def _run_epoch(self, epoch: int, buffers: Dict[int, EpochBuffer]):
        try:
            self._ensure_ingested(epoch)
            t0 = timeit.default_timer()
            if self.stats is not None:
                for _ in range(self.plan.num_reducers):
                    self.stats.reduce_start(epoch)
            key = perm.make_key(self.plan.num_rows, self.seed, epoch)
            n_local = self._packed.shape[0]
            if self.world == 1:
                # K7 on the host: every row is local, so the epoch is produced in
                # destination order - reducer chunk by reducer chunk, round-robin over
                # the trainers - by pulling each position's source row through the
                # inverse permutation (same bytes as a scatter, sequential writes).
                # A chunk is handed out the moment its range is complete (reference
                # dataset.py:133-139: start on the first finished reducer output).
                self._gather_in_chunk_order(key, buffers)
            elif self.C is not None and n_local:
                # native path: pi_e on the C++ worker pool, GIL released
                words = list(key.as_words())
                T = self.plan.num_trainers
                trainer = np.empty(n_local, dtype=np.int32)
                slot = np.empty(n_local, dtype=np.int64)
                self.C.host_perm_positions(self._host_pool, words, self.plan.num_rows, T,
                                           self._offset, n_local, trainer.ctypes.data,
                                           slot.ctypes.data)
                self._exchange(trainer, slot, buffers[self.rank])
            else:
                gidx = np.arange(self._offset, self._offset + n_local, dtype=np.uint64)
                pos = perm.permute(gidx, key)
                trainer, slot = self.plan.position_to_trainer(pos)
                self._exchange(trainer, slot, buffers[self.rank])
            dur = timeit.default_timer() - t0
            if self.stats is not None:
                for _ in range(self.plan.num_reducers):
                    self.stats.reduce_done(epoch, dur / self.plan.num_reducers)
            for buf in buffers.values():
                buf.mark_ready()
        except BaseException as e:  # surface in the consumer, not the pool
            for buf in buffers.values():
                buf.mark_ready(e)

    def _gather_in_chunk_order(self, key, buffers: Dict[int, EpochBuffer]):
        plan, pitch = self.plan, self.layout.row_pitch
        n_local = self._packed.shape[0]
        words = list(key.as_words())
        chunks = {t: plan.trainer_chunks(t) for t in buffers}
        for c in range(max((len(v) for v in chunks.values()), default=0)):
            for t, buf in buffers.items():
                if c >= len(chunks[t]):
                    continue
                a, b = chunks[t][c]
                start = plan.trainer_range(t)[0]
                if b > a and n_local:
                    if self.C is not None:
                        self.C.host_gather_rows(self._host_pool, words, self._packed.ctypes.data,
                                                pitch, self._offset, n_local, start + a, start + b,
                                                buf.data.ctypes.data + a * pitch)
                    else:
                        pos = np.arange(start + a, start + b, dtype=np.uint64)
                        src = perm.inverse(pos, key).astype(np.int64) - self._offset
                        buf.data[a:b] = self._packed[src]
                buf.mark_rows_ready(b)

    def _exchange(self, trainer: np.ndarray, slot: np.ndarray, buf: EpochBuffer):
        """Row exchange over gloo: the host analogue of the NVLink scatter."""
        import torch
        import torch.distributed as dist
        pitch = self.layout.row_pitch
        order = np.argsort(trainer, kind="stable")
        counts = np.bincount(trainer, minlength=self.world).astype(np.int64)
        send_rows = torch.from_numpy(np.ascontiguousarray(self._packed[order]))
        send_slots = torch.from_numpy(np.ascontiguousarray(slot[order]))
        in_counts = torch.from_numpy(counts)
        out_counts = torch.empty_like(in_counts)
        dist.all_to_all_single(out_counts, in_counts, group=self.pg)
        recv_n = int(out_counts.sum())
        recv_rows = torch.empty((recv_n, pitch), dtype=torch.uint8)
        recv_slots = torch.empty(recv_n, dtype=torch.int64)
        osz, isz = out_counts.tolist(), counts.tolist()
        dist.all_to_all_single(recv_rows, send_rows, osz, isz, group=self.pg)
        dist.all_to_all_single(recv_slots, send_slots, osz, isz, group=self.pg)
        buf.data[recv_slots.numpy()] = recv_rows.numpy()

    # -- lifecycle ----------------------------------------------------------
    def release_epoch(self, epoch: int, buffers: Dict[int, EpochBuffer]):
        for b in buffers.values():
            if b.data is not None:
                self._bytes_in_flight -= getattr(b.data, "nbytes", 0)

    def bytes_in_use(self) -> int:
        base = self._packed.nbytes if self._packed is not None else 0
        return base + max(0, self._bytes_in_flight)

    def quiesce(self):
        pass

    def close(self):
        if not self._closed:
            self._closed = True
            from ray_shuffling_data_loader_b200 import stats as stats_mod
            stats_mod.unregister_bytes_used_source(self._bytes_fn)
            self._pool.shutdown(wait=True)
            self._host_pool = None
            self._free.clear()
            self._packed = None
