"""CUDA shuffle engine: HBM-resident columnar table, epoch ring, P2P scatter.

One instance per process (= per GPU). It replaces, for the GPU path, everything
the reference runs as Ray tasks per epoch (``shuffle_map``/``shuffle_reduce``,
reference ``shuffle.py:129-200``) and everything Ray's object store does for it:

* **ingest** - the rows this rank owns are decoded once (``runtime/ingest.py``),
  row group by row group, each decode thread staging through its own pinned slot
  into the HBM-resident columnar table with ``cudaMemcpyAsync`` on a copy stream
  (kernel-table rows K9/K10; decode, page-locking and H2D overlap).
  ``resident="hbm"`` keeps the table on the device for every epoch;
  ``resident="host"`` keeps it in pinned memory and re-streams it chunk by chunk
  each epoch (double-buffered H2D that overlaps the scatter kernel);
  ``resident="disk"`` keeps nothing and re-decodes the row groups every epoch
  through a bounded pinned ring (tables larger than host memory - the reference's
  only mode, ``shuffle.py:151``).
* **epoch ring** (K7, K11) - ``max_concurrent_epochs`` destination slots per local
  trainer inside one arena that every peer maps (VMM symmetric allocation, or
  ``cudaMalloc`` + CUDA IPC), plus epoch-tagged signal words in the same arena:
  ``produced[slot][pass][src_rank]`` (written by each source after each
  destination-chunk pass of its scatter) and ``consumed[trainer]`` (written on the
  trainer's stream when it has finished an epoch - the device analogue of
  ``task_done`` + ``queue.join()``).
* **exchange** - one fused kernel launch per (epoch, destination-chunk pass[,
  source chunk]) pushes every row to its final ``(trainer, slot)`` in local or peer
  HBM (``csrc/shuffle_kernels.cu``); a reducer chunk is consumable as soon as the
  pass that delivers it has completed on every source (reference
  ``dataset.py:133-139``: start on the first finished reducer output).
  ``exchange="nccl"`` swaps in the NCCL ``all_to_all_single`` baseline
  (``parallel/nccl_baseline.py``).

Every wait on a flag has a timeout and raises a clear error instead of hanging
(the reference deadlocks on a dead trainer).
"""
from __future__ import annotations

import ctypes
import os
import threading
import timeit
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.ops import perm
from ray_shuffling_data_loader_b200.ops.plan import ShufflePlan
from ray_shuffling_data_loader_b200.parallel import bootstrap
from ray_shuffling_data_loader_b200.runtime import ingest
from ray_shuffling_data_loader_b200.runtime.chunks import EpochBuffer
from ray_shuffling_data_loader_b200 import stats as stats_mod
from ray_shuffling_data_loader_b200.utils import trace


def load_native():
    """Import the sm_100a extension; on a GPU box a missing build is fatal
    (never fall back silently to a slow path)."""
    try:
        from ray_shuffling_data_loader_b200 import _C
        return _C
    except ImportError as e:
        raise RuntimeError(
            "the native extension ray_shuffling_data_loader_b200._C is not built; "
            "run `python -m ray_shuffling_data_loader_b200._build`") from e


_ROW_PAD = 256          # source columns are padded to this many rows (TMA tiles)
_ALIGN = 256
_FIELD_DTYPE = np.dtype([("src", "<u8"), ("src_code", "<u4"), ("dst_code", "<u4"),
                         ("dst_off", "<u4"), ("width", "<u4")])


# mode-4 conversion kinds (csrc/shuffle_kernels.cu): 8-byte source -> 4-byte field
_KINDS_8_TO_4 = {pair: kind for kind, pair in enumerate(L.TMA_CASTS_8_TO_4)}


def _align(x: int, a: int = _ALIGN) -> int:
    return (x + a - 1) // a * a


class _CudaView:
    """Minimal ``__cuda_array_interface__`` carrier for zero-copy torch views."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str = "|u1"):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
            "version": 3, "strides": None}


def as_torch(ptr: int, shape: Tuple[int, ...], device_index: int, typestr: str = "|u1"):
    import torch
    if int(np.prod(shape)) == 0:
        dt = {"|u1": torch.uint8, "|i1": torch.int8, "<i2": torch.int16, "<i4": torch.int32,
              "<i8": torch.int64, "<u4": torch.int32, "<f2": torch.float16,
              "<f4": torch.float32, "<f8": torch.float64, "|b1": torch.bool}[typestr]
        return torch.empty(shape, dtype=dt, device=f"cuda:{device_index}")
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=f"cuda:{device_index}")


def pinned_array(C, shape, dtype) -> Tuple[np.ndarray, int]:
    """numpy array over freshly allocated pinned host memory -> (array, ptr)."""
    dtype = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape)) * dtype.itemsize)
    ptr = C.pinned_alloc(nbytes)
    raw = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr))
    arr = raw[:int(np.prod(shape)) * dtype.itemsize].view(dtype).reshape(shape)
    return arr, ptr


def default_chunk_passes(layout: L.RowLayout, world: int, max_chunks: int,
                         read_bytes_per_s: float = 2.0e12, link_bytes_per_s: float = 0.68e12) -> int:
    """How many destination-chunk passes stay hidden under the NVLink-bound stores:
    ``floor(t_link / t_read)`` per source row (see DeviceShuffleEngine.__init__), at most 4
    and at most the reducer chunks per trainer; 1 on a single GPU."""
    if world <= 1:
        return 1
    src_row = sum(L.itemsize(f.src_code) * f.width for f in layout.fields)
    t_read = src_row / read_bytes_per_s
    t_link = layout.row_pitch * (world - 1) / world / link_bytes_per_s
    return max(1, min(4, max_chunks, int(t_link / max(t_read, 1e-30))))


class DeviceShuffleEngine:
    device = "cuda"

    def __init__(self, filenames: Sequence[str], plan_args: dict, layout_fn,
                 seed: int, rank: int = 0, world: int = 1, stats_collector=None,
                 max_concurrent_epochs: int = 2, num_threads: Optional[int] = None,
                 resident: str = "hbm", stream_chunk_rows: Optional[int] = None,
                 exchange: str = "p2p", wait_mode: str = "stream",
                 flag_timeout_s: float = 300.0, index=None,
                 device_index: Optional[int] = None, grid: Optional[int] = None,
                 process_group=None, force_generic: bool = False,
                 use_tensor_map: bool = True, peer_alloc: Optional[str] = None,
                 backpressure: Optional[str] = None, numa_bind: Optional[bool] = None,
                 tmap_mode: Optional[int] = None, sched: Optional[int] = None,
                 exchange_group=None, chunk_passes: Optional[int] = None,
                 shuffle_priority: str = "low", tail_fields: bool = True):
        import torch
        self.C = load_native()
        self.torch = torch
        if device_index is None:
            device_index = torch.cuda.current_device()
        self.device_index = device_index
        torch.cuda.set_device(device_index)
        self.C.set_device(device_index)
        cc = self.C.compute_capability(device_index)
        if cc[0] < 10:
            raise RuntimeError(f"sm_100a kernels need a Blackwell GPU, found sm_{cc[0]}{cc[1]}")
        # Put this rank's host side (decode threads, staging copies, pinned
        # buffers) on the socket its GPU is attached to before anything is
        # allocated. Only in one-process-per-GPU mode, and never widening the
        # CPU set the process was given. RSDL_NUMA_BIND=0 disables it.
        if numa_bind is None:
            numa_bind = world > 1 and os.environ.get("RSDL_NUMA_BIND", "1") != "0"
        self.numa_node = self.C.gpu_numa_node(device_index)
        self.numa_cpus = self.C.bind_thread_to_numa_node(self.numa_node) if numa_bind else 0
        self.index = index or ingest.scan_files(filenames)
        self.plan = ShufflePlan(num_rows=self.index.num_rows, **plan_args)
        if world > 1 and self.plan.num_trainers != world:
            raise ValueError("distributed mode needs num_trainers == world size")
        if self.plan.num_trainers > self.C.MAX_TRAINERS:
            raise ValueError(f"at most {self.C.MAX_TRAINERS} trainers are supported")
        self.layout: L.RowLayout = layout_fn(self.index.schema)
        self.seed = int(seed)
        self.rank, self.world, self.pg = rank, world, process_group
        # private communicator of the NCCL baseline exchange (runtime/engine.py)
        self.exchange_pg = exchange_group if exchange_group is not None else process_group
        self.stats = stats_collector
        self.window = max(1, int(max_concurrent_epochs))
        # decode / staging threads: this rank's share of the CPUs it may run on (after
        # NUMA binding), at most 8. Measured on the 128-CPU box (tools/ingest_bench.py,
        # 3.2 GB table in 25 row groups): 8 threads 1.32 s, 16 threads 1.5 - 1.6 s,
        # 32 threads 2.2 - 2.5 s - Arrow's decode buffers and the per-thread pinned
        # slots page-fault against each other beyond that.
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 2
        share = avail // max(1, world) if not self.numa_cpus else avail // max(1, (world + 1) // 2)
        self.num_threads = num_threads or max(2, min(8, share))
        if resident not in ("hbm", "host", "disk"):
            raise ValueError("resident must be 'hbm', 'host' or 'disk'")
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        if wait_mode not in ("host", "stream"):
            raise ValueError("wait_mode must be 'host' or 'stream'")
        self.resident, self.exchange, self.wait_mode = resident, exchange, wait_mode
        self.flag_timeout_s = flag_timeout_s
        self.force_generic = force_generic
        self.fold_tail_fields = bool(tail_fields)     # False: trailing scalars go to the generic kernel
        self.use_tensor_map = use_tensor_map
        # Slot reuse gate (the device analogue of the queue actor's epoch window,
        # reference batch_queue.py:406-418): "stream" enqueues a wait kernel on
        # the shuffle stream so epoch e+window's scatter starts the moment the
        # last trainer's consumed flag lands - no host poll + launch latency
        # between back-to-back shuffles; "host" polls from the driver thread.
        self.backpressure = backpressure or os.environ.get("RSDL_BACKPRESSURE", "stream")
        if self.backpressure not in ("stream", "host"):
            raise ValueError("backpressure must be 'stream' or 'host'")
        if exchange == "nccl":
            self.backpressure = "host"      # collectives are enqueued by torch
        self.peer_alloc = peer_alloc or os.environ.get("RSDL_PEER_ALLOC", "symm")
        if self.peer_alloc not in ("symm", "ipc"):
            raise ValueError("peer_alloc must be 'symm' or 'ipc'")
        # Source tile loads: 0 = 1-D bulk copies (cp.async.bulk, fastest measured:
        # 512-byte contiguous DRAM reads), 1 = tensor-map boxes with 128B swizzle,
        # 2 = one dense tensor-map box per tile. See profiles/README.md.
        # ``sched``: producer schedule of the TMA kernel (-1 = chosen per layout by
        # the binding). Both are constructor options (recorded by bench.py); the
        # RSDL_* environment variables only provide defaults for A/B runs.
        self.tmap_mode = int(os.environ.get("RSDL_TMAP_MODE", "0")) if tmap_mode is None \
            else int(tmap_mode)
        self.sched = int(os.environ.get("RSDL_SCHED", "-1")) if sched is None else int(sched)
        self.local_trainers: List[int] = ([rank] if world > 1
                                          else list(range(self.plan.num_trainers)))
        # K7 - destination-chunk passes. A uniform shuffle makes every destination
        # chunk depend on every source row, so "chunk 0 is ready before the epoch
        # is" is only possible if the sources deliver chunk 0's rows first: the
        # epoch's scatter is issued as ``chunk_passes`` launches, pass k delivering
        # only the rows whose slot falls in the k-th slice of every trainer's
        # buffer, each followed by its own produced flag. Cost: the source table is
        # re-read (and re-indexed) once per pass, which hides under the NVLink-bound
        # stores when world > 1 (default there: as many passes as stay hidden, at most
        # 4 and never more than the reducer chunks per trainer) and is a real cost on
        # one GPU (default 1). Streaming modes and the NCCL baseline use one pass.
        max_chunks = max(self.plan.reducers_of_trainer(t) for t in range(self.plan.num_trainers))
        if chunk_passes is None:
            # A pass re-reads the source at the kernel's load-side rate: ~3.0 TB/s alone
            # (1.06 ms per pass of the 3.2 GB table however few rows it delivers), about
            # 2.0 TB/s while its stores also fight for the link (N = 8, 3 passes: 1.53 ms
            # per pass, epoch 4.62 instead of 4.14 ms - profiles/README.md round 2). The
            # epoch as a whole is bound by NVLink egress (~0.68 TB/s). Passes are free
            # only while a pass's share of the egress takes longer than its re-read, so
            # the default is floor(t_link / t_read) with the pessimistic read rate:
            # 2 passes at N >= 4, 1 at N = 2 for 256-byte f32 rows.
            chunk_passes = default_chunk_passes(self.layout, world, max_chunks)
        if resident != "hbm" or exchange == "nccl":
            chunk_passes = 1
        self.chunk_passes = max(1, min(int(chunk_passes), 64, max(1, self.plan.max_trainer_rows)))
        from ray_shuffling_data_loader_b200.ops.plan import balanced_split
        self.pass_bounds = balanced_split(self.plan.max_trainer_rows, self.chunk_passes)
        self.sm_count = self.C.sm_count(device_index)
        self.grid_override = grid
        self.launches = 0                 # kernels launched by this engine
        self.scatter_launches = 0         # ... of which scatter kernels (any variant)
        self.h2d_bytes_enqueued = 0       # resident="host": bytes handed to cudaMemcpyAsync
        self.epochs_started = 0           # epochs whose shuffle has been enqueued
        self.first_epoch: Optional[int] = None
        self._started_cv = threading.Condition()
        self._lock = threading.Lock()
        self._closed = False
        self._ingested = False
        self._events: Dict[int, Tuple[int, int]] = {}
        self._epoch_kernel_ms: Dict[int, float] = {}
        self._first_pass_events: Dict[int, Tuple[int, int]] = {}
        self._first_pass_ms: Dict[int, float] = {}

        lo, hi = self.plan.source_range(rank, world)
        self.src_lo, self.n_local = lo, hi - lo
        if resident == "disk":
            # the streaming unit is a Parquet row group (re-decoded every epoch):
            # staging buffers hold the largest one this rank owns
            mine = [g for g in self.index.row_groups
                    if g.global_start < hi and g.global_start + g.num_rows > lo and g.num_rows]
            self._disk_groups = mine
            self.chunk_rows = max([min(hi, g.global_start + g.num_rows) - max(lo, g.global_start)
                                   for g in mine] or [1])
        else:
            self.chunk_rows = (self.n_local if resident == "hbm" else
                               max(1, min(self.n_local or 1,
                                          stream_chunk_rows or self.plan.batch_size)))
        if resident != "hbm":
            self.chunk_rows = _align(self.chunk_rows, self.C.TILE_ROWS)

        prio_lo, prio_hi = self.C.stream_priority_range()
        # The shuffle must not starve training kernels: lowest priority stream by
        # default. Caveat (measured, profiles/README.md round 2): the scatter is a
        # persistent kernel of one 512-thread CTA per SM; while a higher-priority
        # stream keeps every SM's 2048 thread slots busy with back-to-back kernels,
        # those CTAs wait for a slot and the epoch starts late. ``"high"`` lets the
        # scatter claim its slots first (a trainer whose kernels need more than
        # ~85 KB of shared memory per SM should keep "low").
        if shuffle_priority not in ("low", "high"):
            raise ValueError("shuffle_priority must be 'low' or 'high'")
        self.shuffle_priority = shuffle_priority
        self.shuffle_stream = self.C.stream_create(prio_lo if shuffle_priority == "low"
                                                   else prio_hi)
        self.copy_stream = self.C.stream_create(prio_lo)
        self.poller = self.C.FlagPoller()
        self.host_pool = self.C.HostPool(self.num_threads)
        self._alloc_dest_arena()
        self._setup_sources()
        self._bytes_fn = self.bytes_in_use
        stats_mod.register_bytes_used_source(self._bytes_fn)
        stats_mod.note_bytes_in_use(self.bytes_in_use())

    # ------------------------------------------------------------------
    # memory
    # ------------------------------------------------------------------
    def _alloc_dest_arena(self):
        C, plan = self.C, self.plan
        T = plan.num_trainers
        self.slot_rows = plan.max_trainer_rows
        self.slot_bytes = _align(max(1, self.slot_rows) * self.layout.row_pitch)
        nloc = len(self.local_trainers)
        # header: produced[W][passes][world] | consumed[T] | error[4]  (uint32 words)
        self.off_produced = 0
        self.off_consumed = _align(self.window * self.chunk_passes * self.world * 4, 128)
        self.off_error = self.off_consumed + _align(T * 4, 128)
        self.header_bytes = _align(self.off_error + 16, 4096)
        self.arena_bytes = self.header_bytes + self.window * nloc * self.slot_bytes
        self._symm = None
        self._opened: List[int] = []
        if self.world > 1 and self.peer_alloc == "symm":
            # VMM-backed symmetric allocation (cuMemCreate + fd exchange, 2 MB
            # pages, explicit peer access): measured 30x faster for the random row
            # scatter than a legacy cudaIpc mapping of a cudaMalloc arena (see
            # profiles/README.md, "NVLink").
            try:
                self._alloc_symmetric()
                return
            except Exception as e:  # pragma: no cover - depends on the driver stack
                import warnings
                warnings.warn(f"symmetric-memory arena unavailable ({e}); "
                              "falling back to legacy CUDA IPC")
        self.arena = C.device_malloc(self.arena_bytes)
        # Zero once: flags start at 0 and row padding stays deterministic.
        C.device_memset_async(self.arena, 0, self.arena_bytes, self.shuffle_stream)
        C.stream_synchronize(self.shuffle_stream)
        # peer mapping (CUDA IPC) - rank r's arena base as seen from this process
        self.peer_base: List[int] = [self.arena] * self.world
        if self.world > 1:
            handle = C.ipc_get_handle(self.arena)
            handles = bootstrap.all_gather_object(
                (self.rank, self.device_index, handle, self.arena_bytes), group=self.pg)
            for r, dev, h, nbytes in handles:
                if r == self.rank:
                    continue
                if nbytes != self.arena_bytes:
                    raise RuntimeError("ranks disagree on the arena size")
                base = C.ipc_open_handle(h)
                self._opened.append(base)
                self.peer_base[r] = base
            bootstrap.barrier(self.pg)

    def _alloc_symmetric(self):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        torch = self.torch
        group = self.pg if self.pg is not None else dist.group.WORLD
        t = symm_mem.empty((self.arena_bytes,), dtype=torch.uint8,
                           device=torch.device("cuda", self.device_index))
        hdl = symm_mem.rendezvous(t, group)
        t.zero_()
        torch.cuda.synchronize(self.device_index)
        hdl.barrier()
        self._symm = (t, hdl)
        self.arena = t.data_ptr()
        self.peer_base = [int(p) for p in hdl.buffer_ptrs]
        if self.peer_base[self.rank] != self.arena:
            self.peer_base[self.rank] = self.arena

    def _owner(self, trainer: int) -> Tuple[int, int]:
        """(owning rank, index among that rank's local trainers)."""
        return (trainer, 0) if self.world > 1 else (0, trainer)

    def _slot_ptr(self, slot: int, trainer: int) -> int:
        owner, j = self._owner(trainer)
        nloc = 1 if self.world > 1 else self.plan.num_trainers
        return (self.peer_base[owner] + self.header_bytes
                + (slot * nloc + j) * self.slot_bytes)

    def _produced_ptr(self, on_rank: int, slot: int, src_rank: int, pass_idx: int = -1) -> int:
        """Flag "source ``src_rank`` delivered pass ``pass_idx`` (default: the last)
        of the epoch in ring slot ``slot``" in rank ``on_rank``'s arena; the
        ``world`` flags of one pass are contiguous (one wait kernel covers them)."""
        k = self.chunk_passes - 1 if pass_idx < 0 else pass_idx
        return (self.peer_base[on_rank] + self.off_produced
                + ((slot * self.chunk_passes + k) * self.world + src_rank) * 4)

    def pass_of_row(self, row_stop: int) -> int:
        """Index of the pass after which rows ``[0, row_stop)`` of a trainer's epoch
        buffer are complete."""
        last = max(0, row_stop - 1)
        for k, (lo, hi) in enumerate(self.pass_bounds):
            if last < hi:
                return k
        return self.chunk_passes - 1

    def _consumed_ptr(self, on_rank: int, trainer: int) -> int:
        return self.peer_base[on_rank] + self.off_consumed + trainer * 4

    def _setup_sources(self):
        """Device storage for the source columns (whole table or 2 staging
        chunks) and the per-buffer column-pointer / field-descriptor arrays."""
        C, lay = self.C, self.layout
        self.src_fields = list(lay.fields)
        rows_buf = _align(max(1, self.chunk_rows), _ROW_PAD)
        self.col_bytes = [_align(rows_buf * L.itemsize(f.src_code) * f.width)
                          for f in self.src_fields]
        self.num_src_bufs = 1 if self.resident == "hbm" else 2
        per_buf = sum(self.col_bytes)
        self.src_arena_bytes = per_buf * self.num_src_bufs + 4096
        self.src_arena = C.device_malloc(self.src_arena_bytes)
        self.src_col_ptrs: List[List[int]] = []
        for b in range(self.num_src_bufs):
            ptrs, off = [], self.src_arena + b * per_buf
            for nb in self.col_bytes:
                ptrs.append(off)
                off += nb
            self.src_col_ptrs.append(ptrs)
        self._plan_kernels()
        # descriptor tables live in a small device buffer, uploaded once
        tables = []
        for b in range(self.num_src_bufs):
            ptrs = self.src_col_ptrs[b]
            fast_cols = np.array([ptrs[i] for i in self.fast_field_idx], dtype=np.uint64)
            runs = []
            for idxs, _lo, _hi in self.generic_runs:
                gen = np.zeros(len(idxs), dtype=_FIELD_DTYPE)
                for k, i in enumerate(idxs):
                    f = self.src_fields[i]
                    gen[k] = (ptrs[i], f.src_code, f.dst_code, f.offset, f.width)
                runs.append(gen)
            tables.append((fast_cols, runs))
        kinds = np.array(self.fast_kinds or [0], dtype=np.uint8)
        blob = b"".join(_pad(a.tobytes()) + b"".join(_pad(g.tobytes()) for g in runs)
                        for a, runs in tables) + _pad(kinds.tobytes())
        self.desc_arena = C.device_malloc(max(256, len(blob)))
        self._desc_host, self._desc_host_ptr = pinned_array(C, (max(256, len(blob)),), np.uint8)
        self._desc_host[:len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        C.memcpy_async(self.desc_arena, self._desc_host_ptr, len(blob), C.H2D, self.shuffle_stream)
        C.stream_synchronize(self.shuffle_stream)
        self.fast_cols_dev, self.generic_fields_dev = [], []
        off = self.desc_arena
        for a, runs in tables:
            self.fast_cols_dev.append(off)
            off += len(_pad(a.tobytes()))
            run_ptrs = []
            for g in runs:
                run_ptrs.append(off)
                off += len(_pad(g.tobytes()))
            self.generic_fields_dev.append(run_ptrs)
        self.fast_kinds_dev = off           # per-column conversion kinds (mode 4)
        if self.resident != "hbm":
            self.h2d_done = [C.event_create(False) for _ in range(self.num_src_bufs)]
            self.buf_free = [C.event_create(False) for _ in range(self.num_src_bufs)]
            self._buf_used = [False] * self.num_src_bufs

    _WIDE_MIN = 16      # list columns at least this wide use the row-major copy kernel

    def _plan_kernels(self):
        """Split the layout's fields between the three scatter kernels
        (csrc/shuffle_kernels.cu): the TMA fast kernel (a dense prefix of 4-byte
        scalar columns), the wide kernel (list-valued columns, already row-major)
        and the generic kernel (everything else, one launch per contiguous run of
        its fields so launches never write each other's bytes)."""
        lay = self.layout
        fields = self.src_fields
        self.fast_mode = -1
        self.fast_kinds: List[int] = []
        self.fast_write_end = 0
        self.fast_field_idx: List[int] = []
        self.wide_field_idx: List[int] = []
        nfast, fast_ranges = 0, []
        if not self.force_generic and fields:
            nfast, fast_ranges, mode = self._fast_prefix()
            if nfast:
                self.fast_mode = mode
                self.fast_field_idx = list(range(nfast))
        rest = list(range(nfast, len(fields)))
        # Tail fields: a few small scalars behind the prefix (a float32 label behind fp8 /
        # bf16 / int64 features, an id behind float features) are folded into the fast
        # kernel's stores instead of costing a second kernel and one more sub-sector write
        # transaction per row (csrc/kernels.h::TailField).
        self.tail_field_idx: List[int] = []
        self.tail_range = (0, 0)
        if nfast and rest and getattr(self, "fold_tail_fields", True) \
                and self._tails_eligible(rest, fast_ranges[0][1]):
            self.tail_field_idx = rest
            hi = lay.scale_offset if self.fast_mode == 2 else lay.row_pitch
            self.tail_range = (fast_ranges[0][1], hi)
            fast_ranges = [(0, hi)] + list(fast_ranges[1:])
            rest = []
        if not self.force_generic:
            self.wide_field_idx = [i for i in rest if fields[i].width >= self._WIDE_MIN]
        generic = [i for i in rest if i not in self.wide_field_idx]
        # contiguous runs of generic fields (by offset) between fast / wide ranges
        taken = list(fast_ranges) if nfast else []
        taken += [(fields[i].offset, fields[i].offset + fields[i].dst_bytes)
                  for i in self.wide_field_idx]
        generic.sort(key=lambda i: fields[i].offset)
        runs: List[Tuple[List[int], int, int]] = []
        for i in generic:
            lo = fields[i].offset // 4 * 4
            hi = _align(fields[i].offset + fields[i].dst_bytes, 4)
            if runs:
                idxs, rlo, rhi = runs[-1]
                blocked = any(a < hi and b > rhi for a, b in taken)   # something in between
                if not blocked:
                    runs[-1] = (idxs + [i], rlo, max(rhi, hi))
                    continue
            runs.append(([i], lo, hi))
        if not nfast and not self.wide_field_idx and runs:
            # pure generic layout: one launch owns the whole row (padding included)
            runs = [(sum((r[0] for r in runs), []), 0, lay.row_pitch)]
        for idxs, lo, hi in runs:
            for a, b in taken:
                if a < hi and b > lo:
                    raise ValueError("packed-row fields overlap between kernels; "
                                     "reorder feature columns so same-typed scalars are adjacent")
        self.generic_runs = runs
        self.generic_field_idx = [i for r in runs for i in r[0]]

    def _tails_eligible(self, idxs: List[int], fast_end: int) -> bool:
        """May the fields ``idxs`` (everything behind the fast prefix) ride the fast
        kernel as tail fields? At most 4 scalars, 4- or 8-byte destinations, only bit
        copies and the three 8 -> 4 byte conversions, inside [fast_end, pitch) - and, with
        the fp8 scale bytes, in front of them."""
        lay, fields = self.layout, self.src_fields
        if len(idxs) > self.C_MAX_TAIL or fast_end % 16:
            return False
        hi = lay.scale_offset if self.fast_mode == 2 else lay.row_pitch
        for i in idxs:
            f = fields[i]
            ssz, dsz = L.itemsize(f.src_code), L.itemsize(f.dst_code)
            if f.width != 1 or dsz not in (4, 8) or ssz not in (4, 8):
                return False
            same = f.dst_code == f.src_code
            if not (same or (ssz == 8 and (f.src_code, f.dst_code) in _KINDS_8_TO_4)):
                return False
            if f.offset < fast_end or f.offset + dsz > hi or f.offset % dsz:
                return False
        return True

    C_MAX_TAIL = 4      # RSDL_MAX_TAIL_FIELDS (csrc/kernels.h)

    def _fast_prefix(self):
        """(number of leading fields the TMA kernel takes, byte ranges it writes,
        mode) - (0, [], -1) when the layout does not start with such a prefix."""
        lay, fields = self.layout, self.src_fields
        first = fields[0]
        ssz = L.itemsize(first.src_code)
        if first.width != 1 or ssz not in (4, 8) or first.offset != 0:
            return 0, [], -1
        dsz = L.itemsize(first.dst_code)
        if ssz == 8:
            # 8-byte sources (int64 / float64, the reference's DATA_SPEC): bit copy
            # (mode 3) or per-column conversion to a 4-byte destination (mode 4)
            mode = {8: 3, 4: 4}.get(dsz, -1)
            if mode < 0:
                return 0, [], -1

            def same_class(f):
                if L.itemsize(f.src_code) != 8 or L.itemsize(f.dst_code) != dsz:
                    return False
                return (f.dst_code == f.src_code if mode == 3
                        else (f.src_code, f.dst_code) in _KINDS_8_TO_4)
        else:
            if first.dst_code == first.src_code:
                mode = 0
            elif first.src_code == L.DT_F32 and first.dst_code == L.DT_BF16:
                mode = 1
            elif (first.src_code == L.DT_F32 and first.dst_code == L.DT_FP8
                  and lay.scale_offset >= 0):
                mode = 2
            else:
                return 0, [], -1

            def same_class(f):
                return f.src_code == first.src_code and f.dst_code == first.dst_code
        n, off = 0, 0
        for f in fields:
            if f.width != 1 or not same_class(f) or f.offset != off:
                break
            n += 1
            off += dsz
        # The kernel writes whole 16-byte groups. When another field starts inside
        # the prefix's last, partly filled group (e.g. 5 x f32 then an int64 at
        # byte 24), give the odd fields of that group to the generic kernel
        # rather than the whole row (mode 2's blocks of 32 cannot be split).
        per_group = 16 // dsz
        if mode != 2 and n < len(fields) and n % per_group \
                and fields[n].offset < _align(off, 16):
            n -= n % per_group
            off = n * dsz
        if mode == 4:
            self.fast_kinds = [_KINDS_8_TO_4[(f.src_code, f.dst_code)] for f in fields[:n]]
        if n < 4:
            return 0, [], -1       # not worth a TMA launch
        fast_end = _align(off, 16)
        rest = fields[n:]
        if rest:
            rest_lo = min(f.offset for f in rest)
            rest_hi = max(f.offset + f.dst_bytes for f in rest)
            scale_lo = lay.scale_offset if mode == 2 else lay.row_pitch
            if rest_lo < fast_end or (mode == 2 and rest_hi > scale_lo):
                return 0, [], -1
        # with nothing behind the prefix the kernel also zero-fills the row's
        # padding: every 32-byte sector of a row is then written whole
        self.fast_write_end = fast_end if (rest or mode == 2) else lay.row_pitch
        ranges = [(0, self.fast_write_end)]
        if mode == 2:               # plus the UE8M0 scale bytes after everything else
            ranges.append((lay.scale_offset, lay.scale_offset + (n + 31) // 32))
        return n, ranges, mode

    # ------------------------------------------------------------------
    # ingest
    # ------------------------------------------------------------------
    def _ensure_ingested(self, epoch: int):
        with self._lock:
            if self._ingested:
                if self.stats is not None:
                    for _ in range(len(self.index.filenames)):
                        self.stats.map_start(epoch)
                        self.stats.map_done(epoch, 0.0, 0.0)
                return
            C = self.C
            t0 = timeit.default_timer()
            nfiles = len(self.index.filenames)
            if self.resident == "disk":
                # nothing is kept: every epoch re-decodes its row groups (_stream_epoch_disk)
                self._pinned = []
                self._ingested = True
                self.ingest_seconds = 0.0
                self.decode_seconds = 0.0
                return
            if self.stats is not None:
                for _ in range(nfiles):
                    self.stats.map_start(epoch)
            self._pinned: List[int] = []
            if self.resident == "hbm":
                reads = self._ingest_hbm_through_ring()
                self._ingested = True
                self.ingest_seconds = dur = timeit.default_timer() - t0
                if self.stats is not None:
                    for _ in range(nfiles):
                        self.stats.map_done(epoch, dur / nfiles, float(np.mean(reads or [0.0])))
                return

            # resident="host": the whole decoded table stays pinned.
            # One pinned block, bump-allocated: same-shape columns end up with a
            # uniform stride, so a whole chunk moves with ONE cudaMemcpy2DAsync.
            dtypes = {f.name: np.dtype(L.numpy_storage_dtype(f.src_code)) for f in self.src_fields}
            widths = {f.name: f.width for f in self.src_fields}
            total = sum(_align(self.n_local * dtypes[n].itemsize * max(1, widths[n]))
                        for n in dict.fromkeys(f.name for f in self.src_fields)) + 4096
            t_pin = timeit.default_timer()
            host_block, host_ptr = pinned_array(C, (total,), np.uint8)
            self.pinned_alloc_seconds = timeit.default_timer() - t_pin
            self._pinned.append(host_ptr)
            cursor = {"off": 0}
            bump_lock = threading.Lock()

            def alloc(shape, dt):
                dt = np.dtype(dt)
                nbytes = int(np.prod(shape)) * dt.itemsize
                with bump_lock:
                    off = cursor["off"]
                    if off + nbytes > total:        # unexpected shape: private block
                        arr, ptr = pinned_array(C, shape, dt)
                        self._pinned.append(ptr)
                        return arr
                    cursor["off"] = off + _align(nbytes)
                return host_block[off:off + nbytes].view(dt).reshape(shape)

            names = [f.name for f in self.src_fields]
            t_dec = timeit.default_timer()
            table = ingest.load_table(self.index, self.src_lo, self.src_lo + self.n_local,
                                      columns=list(dict.fromkeys(names)),
                                      num_threads=self.num_threads, alloc=alloc,
                                      copy_fn=self._host_copy)
            self.decode_seconds = timeit.default_timer() - t_dec
            self.host_table = table
            self.host_cols = [table.columns[f.name] for f in self.src_fields]
            for f, col in zip(self.src_fields, self.host_cols):
                want = np.dtype(L.numpy_storage_dtype(f.src_code))
                if col.dtype != want:
                    raise TypeError(f"column {f.name}: decoded {col.dtype}, expected {want}")
            self._ingested = True
            dur = timeit.default_timer() - t0
            self.ingest_seconds = dur
            reads = table.read_durations or [0.0]
            if self.stats is not None:
                for _ in range(nfiles):
                    self.stats.map_done(epoch, dur / nfiles, float(np.mean(reads)))

    def _ingest_hbm_through_ring(self) -> List[float]:
        """Cold path for ``resident="hbm"``: Parquet -> HBM without ever pinning the
        whole table. Every decode thread owns ONE pinned staging slot of row-group
        size (allocated by the thread itself, so page-locking overlaps the other
        threads' decoding - cudaHostAlloc costs ~0.27 s per GB and used to be 45 % of
        the ingest); it decodes a row group straight into the slot (Arrow -> numpy
        views over the slot), hands each column slice to the copy engine and only
        waits for that copy when it needs the slot again. Decode, page-locking and
        H2D all overlap; pinned memory in use = threads x one row group."""
        from concurrent.futures import ThreadPoolExecutor
        C = self.C
        lo_all, hi_all = self.src_lo, self.src_lo + self.n_local
        groups = [g for g in self.index.row_groups
                  if g.global_start < hi_all and g.global_start + g.num_rows > lo_all and g.num_rows]
        self.decode_seconds = 0.0
        self.pinned_alloc_seconds = 0.0
        if not groups:
            return []
        max_rows = max(min(hi_all, g.global_start + g.num_rows) - max(lo_all, g.global_start)
                       for g in groups)
        names = list(dict.fromkeys(f.name for f in self.src_fields))
        first = {f.name: f for f in reversed(self.src_fields)}
        host_off, off = {}, 0
        for n in names:
            f = first[n]
            host_off[n] = off
            off += _align(max_rows * L.itemsize(f.src_code) * max(1, f.width))
        slot_bytes = max(off, 256)
        by_name: Dict[str, List[int]] = {}
        for i, f in enumerate(self.src_fields):
            by_name.setdefault(f.name, []).append(i)
        tls = threading.local()
        slots, lock = [], threading.Lock()
        stats_lock = threading.Lock()
        reads: List[float] = []
        t_alloc = [0.0]
        dev_index = self.device_index

        def work(g):
            C.set_device(dev_index)                   # device is per-thread
            slot = getattr(tls, "slot", None)
            if slot is None:
                t0 = timeit.default_timer()
                arr, ptr = pinned_array(C, (slot_bytes,), np.uint8)
                slot = tls.slot = {"arr": arr, "ptr": ptr, "copied": C.event_create(False),
                                   "used": False}
                with lock:
                    slots.append(slot)
                    t_alloc[0] += timeit.default_timer() - t0
            elif slot["used"]:
                C.event_synchronize(slot["copied"])   # previous row group has left the slot
            lo, hi = max(lo_all, g.global_start), min(hi_all, g.global_start + g.num_rows)
            bufs = {}
            for n in names:
                f = first[n]
                dt = np.dtype(L.numpy_storage_dtype(f.src_code))
                cnt = (hi - lo) * max(1, f.width)
                view = slot["arr"][host_off[n]:host_off[n] + cnt * dt.itemsize].view(dt)
                bufs[n] = view if f.width == 1 else view.reshape(hi - lo, f.width)
            table = ingest.load_table(self.index, lo, hi, columns=names, num_threads=1,
                                      prealloc=bufs)
            for n in names:
                if table.columns[n].dtype != bufs[n].dtype:
                    raise TypeError(f"column {n}: decoded {table.columns[n].dtype}, "
                                    f"expected {bufs[n].dtype}")
            with lock:                                 # one enqueuer at a time per stream
                for n in names:
                    for i in by_name[n]:
                        f = self.src_fields[i]
                        isz = L.itemsize(f.src_code) * f.width
                        C.memcpy_async(self.src_col_ptrs[0][i] + (lo - lo_all) * isz,
                                       slot["ptr"] + host_off[n], (hi - lo) * isz, C.H2D,
                                       self.copy_stream)
                C.event_record(slot["copied"], self.copy_stream)
            slot["used"] = True
            with stats_lock:
                reads.extend(table.read_durations or [])

        t_dec = timeit.default_timer()
        workers = max(1, min(self.num_threads, len(groups)))
        with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="ingest") as ex:
            list(ex.map(work, groups))
        C.stream_synchronize(self.copy_stream)
        self.decode_seconds = timeit.default_timer() - t_dec
        self.pinned_alloc_seconds = t_alloc[0] / max(1, len(slots))     # mean per thread
        for slot in slots:
            C.event_destroy(slot["copied"])
            C.pinned_free(slot["ptr"])
        self.host_cols = None
        self.host_table = None
        return reads

    # ------------------------------------------------------------------
    # one epoch
    # ------------------------------------------------------------------
    def _grid(self, work_items: int, ctas_per_sm: int = 1) -> int:
        g = self.grid_override or self.sm_count * ctas_per_sm
        return max(1, min(g, work_items))

    def _launch_chunk(self, key_words, buf: int, n_rows: int, global_offset: int,
                      dst: List[int], slot_range: Optional[Tuple[int, int]] = None):
        C, lay, plan = self.C, self.layout, self.plan
        if n_rows <= 0:
            return
        slot_lo, slot_hi = slot_range if slot_range is not None else (0, (1 << 64) - 1)
        if self.fast_mode >= 0:
            ncols = len(self.fast_field_idx)
            tiles = -(-n_rows // C.fast_tile_rows(self.fast_mode))
            # fast columns are contiguous with one stride: hand the kernel a 2-D
            # tensor map (4 TMA box loads per tile instead of 64 bulk copies)
            ptrs = self.src_col_ptrs[buf]
            stride = self.col_bytes[0]
            uniform = (self.use_tensor_map and self.tmap_mode != 0 and self.fast_mode <= 2 and all(
                ptrs[i] == ptrs[0] + i * stride for i in range(ncols)))
            C.scatter_fast(key=key_words, num_rows=plan.num_rows,
                           num_trainers=plan.num_trainers, cols=self.fast_cols_dev[buf],
                           num_cols=ncols, n_local=n_rows, global_offset=global_offset,
                           row_pitch=lay.row_pitch,
                           scale_offset=max(0, lay.scale_offset), dst=dst,
                           mode=self.fast_mode,
                           grid=self._grid(tiles, C.fast_ctas_per_sm(self.fast_mode)),
                           stream=self.shuffle_stream,
                           col_base=ptrs[0] if uniform else 0,
                           col_stride=stride if uniform else 0,
                           rows_alloc=(stride // 4) if uniform else 0,
                           tmap_mode=self.tmap_mode,
                           kinds=self.fast_kinds_dev if self.fast_mode == 4 else 0,
                           write_end=self.fast_write_end,
                           sched=self.sched, slot_lo=slot_lo, slot_hi=slot_hi,
                           tail=[(ptrs[i], self.src_fields[i].src_code,
                                  self.src_fields[i].dst_code, self.src_fields[i].offset)
                                 for i in self.tail_field_idx],
                           tail_lo=self.tail_range[0], tail_hi=self.tail_range[1])
            self.launches += 1
            self.scatter_launches += 1
        for i in self.wide_field_idx:
            f = self.src_fields[i]
            C.scatter_wide(key=key_words, num_rows=plan.num_rows,
                           num_trainers=plan.num_trainers, src=self.src_col_ptrs[buf][i],
                           width=f.width, src_code=f.src_code, dst_code=f.dst_code,
                           dst_off=f.offset, n_local=n_rows, global_offset=global_offset,
                           row_pitch=lay.row_pitch, dst=dst, grid=self.grid_override or 0,
                           stream=self.shuffle_stream, slot_lo=slot_lo, slot_hi=slot_hi)
            self.launches += 1
            self.scatter_launches += 1
        for (idxs, lo, hi), fields_dev in zip(self.generic_runs, self.generic_fields_dev[buf]):
            C.scatter_generic(key=key_words, num_rows=plan.num_rows,
                              num_trainers=plan.num_trainers, fields=fields_dev,
                              num_fields=len(idxs), n_local=n_rows,
                              global_offset=global_offset, row_pitch=lay.row_pitch,
                              write_lo=lo, write_hi=hi, dst=dst,
                              grid=self.grid_override or 0,   # 0: launcher picks by occupancy
                              stream=self.shuffle_stream, slot_lo=slot_lo, slot_hi=slot_hi)
            self.launches += 1
            self.scatter_launches += 1

    def _signal_produced(self, slot: int, pass_idx: int, epoch: int):
        """publish: produced[slot][pass][this rank] = epoch + 1 on every rank
        (st.release.sys after a system fence, stream-ordered after the pass)."""
        targets = [self._produced_ptr(r, slot, self.rank, pass_idx) for r in range(self.world)]
        self.C.signal_flags(targets, epoch + 1, self.shuffle_stream)
        self.launches += 1

    def start_epoch(self, epoch: int) -> Dict[int, EpochBuffer]:
        """Enqueue the epoch's shuffle on the side stream (non-blocking apart
        from cross-rank back-pressure) and return the destination buffers."""
        if self._closed:
            raise RuntimeError("engine is closed")
        C, plan = self.C, self.plan
        C.set_device(self.device_index)      # driver thread: device is per-thread
        self.torch.cuda.set_device(self.device_index)
        with trace.span("ingest", epoch=epoch):
            self._ensure_ingested(epoch)
        slot = epoch % self.window
        if self.first_epoch is None:
            # Resume (start_epoch = k): the arena's flags start at zero in a fresh
            # process, so the slot-reuse gate only applies to epochs that reuse a
            # slot *this engine* has filled (epoch - first >= window). Flags are
            # monotonic epoch tags, so nothing needs seeding.
            self.first_epoch = epoch
        gate = epoch - self.first_epoch >= self.window
        if gate and self.backpressure == "stream":
            # Back-pressure, stream ordered: every trainer must have released
            # this slot's previous epoch before any source may overwrite it.
            C.wait_flags(self._consumed_ptr(self.rank, 0), plan.num_trainers,
                         epoch - self.window + 1, int(self.flag_timeout_s * 1e9),
                         self.arena + self.off_error, self.shuffle_stream)
            self.launches += 1
        elif gate:
            with trace.span("backpressure_wait", epoch=epoch):
                lag = self._poll(self._consumed_ptr(self.rank, 0), plan.num_trainers,
                                 epoch - self.window + 1, self.flag_timeout_s)
            if lag >= 0:
                raise TimeoutError(
                    f"trainer {lag} did not release epoch {epoch - self.window} within "
                    f"{self.flag_timeout_s}s (stalled or dead consumer)")
        t_start = timeit.default_timer()
        if self.stats is not None:
            for _ in range(plan.num_reducers):
                self.stats.reduce_start(epoch)
        key_words = list(perm.make_key(plan.num_rows, self.seed, epoch).as_words())
        dst = [self._slot_ptr(slot, t) for t in range(plan.num_trainers)]
        ev0, ev1 = C.event_create(True), C.event_create(True)
        C.event_record(ev0, self.shuffle_stream)
        trace.instant("launch_epoch_shuffle", epoch=epoch, slot=slot)
        ev_first = ev_first0 = None
        if self.chunk_passes > 1:
            ev_first0 = C.event_create(True)
            C.event_record(ev_first0, self.shuffle_stream)
        if self.exchange == "nccl":
            from ray_shuffling_data_loader_b200.parallel import nccl_baseline
            nccl_baseline.exchange_epoch(self, key_words, slot)
            self._signal_produced(slot, 0, epoch)
        elif self.resident == "hbm":
            # one launch (set) per destination-chunk pass, each with its own flag
            for k, bounds in enumerate(self.pass_bounds):
                self._launch_chunk(key_words, 0, self.n_local, self.src_lo, dst,
                                   bounds if self.chunk_passes > 1 else None)
                self._signal_produced(slot, k, epoch)
                if k == 0 and self.chunk_passes > 1:
                    ev_first = C.event_create(True)
                    C.event_record(ev_first, self.shuffle_stream)
        elif self.resident == "disk":
            self._stream_epoch_disk(key_words, dst, epoch)
            self._signal_produced(slot, 0, epoch)
        else:
            self._stream_epoch(key_words, dst)
            self._signal_produced(slot, 0, epoch)
        C.event_record(ev1, self.shuffle_stream)
        if ev_first is not None:
            self._first_pass_events[epoch] = (ev_first0, ev_first)
        self._events[epoch] = (ev0, ev1)

        buffers = {}
        for t in self.local_trainers:
            rows = plan.trainer_rows(t)
            data = as_torch(self._slot_ptr(slot, t), (rows, self.layout.row_pitch),
                            self.device_index)
            buffers[t] = EpochBuffer(
                epoch, t, rows, self.layout, data, "cuda",
                wait_fn=self._make_wait(epoch, slot, t_start),
                release_fn=self._make_release(epoch, t))
        with self._started_cv:
            self.epochs_started = epoch + 1
            self._started_cv.notify_all()
        return buffers

    def wait_epochs_started(self, n: int, timeout: float = 60.0) -> bool:
        """Block until the shuffles of epochs ``< n`` have been enqueued (the
        driver thread runs ahead of the consumer by the epoch window)."""
        with self._started_cv:
            return self._started_cv.wait_for(lambda: self.epochs_started >= n, timeout)

    def enqueue_wait_produced(self, epoch: int, stream: Optional[int] = None):
        """Make ``stream`` (default: the current torch stream) wait, on the
        device, until every source rank has delivered ``epoch`` to this rank."""
        if stream is None:
            stream = self.torch.cuda.current_stream().cuda_stream
        self.C.wait_flags(self._produced_ptr(self.rank, epoch % self.window, 0), self.world,
                          epoch + 1, int(self.flag_timeout_s * 1e9),
                          self.arena + self.off_error, stream)
        self.launches += 1

    def _stream_epoch(self, key_words, dst):
        """resident='host': double-buffered H2D of source chunks overlapping the
        scatter kernel of the previous chunk."""
        C = self.C
        k = 0
        for row0 in range(0, self.n_local, self.chunk_rows):
            rows = min(self.chunk_rows, self.n_local - row0)
            b = k % self.num_src_bufs
            if self._buf_used[b]:
                C.stream_wait_event(self.copy_stream, self.buf_free[b])
            self._h2d_chunk(b, row0, rows)
            C.event_record(self.h2d_done[b], self.copy_stream)
            C.stream_wait_event(self.shuffle_stream, self.h2d_done[b])
            self._launch_chunk(key_words, b, rows, self.src_lo + row0, dst)
            C.event_record(self.buf_free[b], self.shuffle_stream)
            self._buf_used[b] = True
            k += 1

    def _stream_epoch_disk(self, key_words, dst, epoch: int):
        """resident='disk' - tables larger than pinned host memory (the reference's
        only mode: it re-reads Parquet every epoch, ``shuffle.py:151``). Row groups
        are decoded by a bounded pool into a ring of pinned staging buffers (decode
        threads + 2 of them, each one row group), copied to one of two
        device staging buffers and scattered; decode of the next row groups, the
        H2D of this one and the scatter of the previous one overlap. Host memory
        in use is bounded by the ring, not by the table."""
        from concurrent.futures import ThreadPoolExecutor
        C = self.C
        groups = self._disk_groups
        if not groups:
            return
        if not hasattr(self, "_disk_ring"):
            depth = max(2, min(self.num_threads, len(groups))) + 2
            per_buf = sum(self.col_bytes)
            self._disk_ring = []
            for _ in range(depth):
                arr, ptr = pinned_array(C, (per_buf,), np.uint8)
                self._pinned.append(ptr)
                self._disk_ring.append({"arr": arr, "ptr": ptr, "copied": C.event_create(False),
                                        "used": False})
            self._disk_pool = ThreadPoolExecutor(max_workers=depth - 2,
                                                 thread_name_prefix="disk-decode")
        ring = self._disk_ring
        host_off, off = {}, 0           # a column's place inside a pinned staging slot
        for f, nb in zip(self.src_fields, self.col_bytes):
            host_off.setdefault(f.name, off)
            off += nb
        names = list(dict.fromkeys(f.name for f in self.src_fields))
        lo_all, hi_all = self.src_lo, self.src_lo + self.n_local
        stats = self.stats
        if stats is not None:
            for _ in range(len(self.index.filenames)):
                stats.map_start(epoch)

        def decode(g, slot):
            """one row group -> pinned staging slot (columns at the device layout's offsets)"""
            if slot["used"]:
                C.event_synchronize(slot["copied"])        # its previous H2D has drained
            lo, hi = max(lo_all, g.global_start), min(hi_all, g.global_start + g.num_rows)
            offs = host_off

            t0 = timeit.default_timer()
            bufs = {}
            for f in self.src_fields:
                if f.name in bufs:
                    continue
                dt = np.dtype(L.numpy_storage_dtype(f.src_code))
                n = (hi - lo) * max(1, f.width)
                view = slot["arr"][offs[f.name]:offs[f.name] + n * dt.itemsize].view(dt)
                bufs[f.name] = view if f.width == 1 else view.reshape(hi - lo, f.width)
            table = ingest.load_table(self.index, lo, hi, columns=names, num_threads=1,
                                      prealloc=bufs)
            return lo, hi - lo, timeit.default_timer() - t0, (table.read_durations or [0.0])[0]

        pending = []
        depth = len(ring)
        it = iter(enumerate(groups))

        def submit_next():
            nxt = next(it, None)
            if nxt is None:
                return
            k, g = nxt
            slot = ring[k % depth]
            pending.append((slot, self._disk_pool.submit(decode, g, slot)))
        for _ in range(depth - 2):
            submit_next()
        k = 0
        read_times = []
        while pending:
            slot, fut = pending.pop(0)
            row0, rows, total_s, read_s = fut.result()
            read_times.append(read_s)
            b = k % self.num_src_bufs
            if self._buf_used[b]:
                C.stream_wait_event(self.copy_stream, self.buf_free[b])
            for i, f in enumerate(self.src_fields):
                isz = L.itemsize(f.src_code) * f.width
                C.memcpy_async(self.src_col_ptrs[b][i], slot["ptr"] + host_off[f.name],
                               rows * isz, C.H2D, self.copy_stream)
                self.h2d_bytes_enqueued += rows * isz
            C.event_record(slot["copied"], self.copy_stream)
            slot["used"] = True
            C.event_record(self.h2d_done[b], self.copy_stream)
            C.stream_wait_event(self.shuffle_stream, self.h2d_done[b])
            self._launch_chunk(key_words, b, rows, row0, dst)
            C.event_record(self.buf_free[b], self.shuffle_stream)
            self._buf_used[b] = True
            k += 1
            submit_next()
        if stats is not None:
            nfiles = len(self.index.filenames)
            for _ in range(nfiles):
                stats.map_done(epoch, 0.0, float(np.mean(read_times or [0.0])))

    def _host_copy(self, dst: np.ndarray, src: np.ndarray):
        """Decoded Arrow buffer -> pinned staging on the C++ worker pool."""
        self.host_pool.parallel_memcpy(dst.ctypes.data, src.ctypes.data, src.nbytes)

    def _h2d_chunk(self, b: int, row0: int, rows: int):
        """Copy rows [row0, row0+rows) of every source column into staging buffer
        ``b``: runs of columns with a uniform host stride go as one 2-D copy."""
        C = self.C
        ptrs, cols, fields = self.src_col_ptrs[b], self.host_cols, self.src_fields
        i, n = 0, len(fields)
        while i < n:
            isz = L.itemsize(fields[i].src_code) * fields[i].width
            base = cols[i].ctypes.data
            j = i + 1
            if j < n:
                hstride = cols[j].ctypes.data - base
                dstride = ptrs[j] - ptrs[i]
                while (j < n and hstride > 0 and dstride > 0
                       and L.itemsize(fields[j].src_code) * fields[j].width == isz
                       and cols[j].ctypes.data - base == (j - i) * hstride
                       and ptrs[j] - ptrs[i] == (j - i) * dstride):
                    j += 1
            if j - i >= 2:
                C.memcpy2d_async(ptrs[i], dstride, base + row0 * isz, hstride, rows * isz,
                                 j - i, C.H2D, self.copy_stream)
            else:
                j = i + 1
                C.memcpy_async(ptrs[i], base + row0 * isz, rows * isz, C.H2D, self.copy_stream)
            self.h2d_bytes_enqueued += rows * isz * (j - i)
            i = j

    def h2d_bytes_per_epoch(self) -> int:
        if self.resident == "hbm":
            return 0
        return int(sum(self.n_local * L.itemsize(f.src_code) * f.width
                       for f in self.src_fields))

    def cancel(self):
        """Make every pending flag wait fail fast (early teardown)."""
        self._cancelled = True

    def _poll(self, ptr: int, count: int, value: int, timeout_s: float) -> int:
        """Host-side flag wait in short native slices so teardown can interrupt
        it; returns -1 when all flags reached ``value`` else the lagging index."""
        deadline = timeit.default_timer() + timeout_s
        while True:
            if getattr(self, "_cancelled", False) or self._closed:
                raise RuntimeError("shuffle engine is shutting down")
            remaining = deadline - timeit.default_timer()
            lag = self.poller.wait(ptr, count, value, max(0.0, min(0.05, remaining)))
            if lag < 0 or remaining <= 0:
                return lag

    def _make_wait(self, epoch: int, slot: int, t_start: float):
        state = {"pass": -1}

        def wait(timeout: Optional[float] = None, row_stop: Optional[int] = None):
            """Wait until rows ``[0, row_stop)`` of the buffer (default: all of it)
            have been delivered by every source - i.e. for the produced flags of
            the destination-chunk pass that completes them (K7)."""
            need = self.chunk_passes - 1 if row_stop is None else self.pass_of_row(row_stop)
            if need <= state["pass"]:
                return
            self.C.set_device(self.device_index)
            flags = self._produced_ptr(self.rank, slot, 0, need)
            if self.wait_mode == "stream":
                # Device-side wait on the consumer's stream: no host sync. A wait
                # kernel that gave up (dead peer) only sets the error word, so
                # look at it once per epoch - an unordered 4-byte read on the
                # poller's stream - and fail within an epoch instead of feeding
                # the trainer a half-written buffer.
                if epoch > 0 and state["pass"] < 0:
                    self.check_error()
                stream = self.torch.cuda.current_stream().cuda_stream
                self.C.wait_flags(flags, self.world, epoch + 1, int(self.flag_timeout_s * 1e9),
                                  self.arena + self.off_error, stream)
                self.launches += 1
            else:
                limit = self.flag_timeout_s if timeout is None else timeout
                with trace.span("wait_epoch_produced", epoch=epoch):
                    lag = self._poll(flags, self.world, epoch + 1, limit)
                if lag >= 0:
                    self.check_error()      # a device-side wait that gave up explains it
                    raise TimeoutError(
                        f"source rank {lag} did not deliver epoch {epoch} (pass {need}) "
                        f"within {limit}s")
            state["pass"] = need
            if need == self.chunk_passes - 1:
                self._epoch_done_stats(epoch, t_start)
        return wait

    def _epoch_done_stats(self, epoch: int, t_start: float):
        if epoch not in self._events:
            return
        ms = self._resolve_kernel_ms(epoch)
        if self.stats is not None:
            dur = (ms / 1e3) if ms is not None else (timeit.default_timer() - t_start)
            R = self.plan.num_reducers
            for _ in range(R):
                self.stats.reduce_done(epoch, dur / R)
            if ms is not None:
                remote = self.n_local * self.layout.row_pitch * (self.world - 1) // max(1, self.world)
                self.stats.exchange_done(epoch, remote, ms / 1e3)

    def _resolve_kernel_ms(self, epoch: int) -> Optional[float]:
        """CUDA-event time of the epoch's shuffle on the shuffle stream, once
        its end event has completed (a stream-mode wait returns to the host long
        before that, so the events are kept until somebody asks again)."""
        ev = self._events.get(epoch)
        if ev is None:
            return self._epoch_kernel_ms.get(epoch)
        try:
            if not self.C.event_query(ev[1]):
                return None
            ms = self.C.event_elapsed_ms(ev[0], ev[1])
        except Exception:
            return None
        self._events.pop(epoch, None)
        self.C.event_destroy(ev[0])
        self.C.event_destroy(ev[1])
        self._epoch_kernel_ms[epoch] = ms
        return ms

    def first_pass_ms(self, epoch: int) -> Optional[float]:
        """Device time from the start of the epoch's shuffle to the end of its
        first destination-chunk pass (= when chunk 0 becomes consumable); ``None``
        with a single pass or before the pass has finished."""
        if epoch in self._first_pass_ms:
            return self._first_pass_ms[epoch]
        ev = self._first_pass_events.get(epoch)
        if ev is None:
            return None
        try:
            if not self.C.event_query(ev[1]):
                return None
            ms = self.C.event_elapsed_ms(ev[0], ev[1])
        except Exception:
            return None
        self._first_pass_events.pop(epoch)
        self.C.event_destroy(ev[0])
        self.C.event_destroy(ev[1])
        self._first_pass_ms[epoch] = ms
        return ms

    def epoch_kernel_ms(self, epoch: int) -> Optional[float]:
        if epoch in self._epoch_kernel_ms:
            return self._epoch_kernel_ms[epoch]
        return self._resolve_kernel_ms(epoch)

    def source_column_tensors(self):
        """Typed torch views of the HBM-resident source columns (``resident="hbm"``
        after ingest): used by tests / bench to compute ground-truth sums with
        plain torch, independent of the shuffle kernels."""
        if self.resident != "hbm" or not self._ingested:
            raise RuntimeError("source columns are only device resident with resident='hbm'")
        out = []
        for f, ptr in zip(self.src_fields, self.src_col_ptrs[0]):
            typestr = np.dtype(L.numpy_storage_dtype(f.src_code)).str
            shape = (self.n_local,) if f.width == 1 else (self.n_local, f.width)
            out.append((f, as_torch(ptr, shape, self.device_index, typestr)))
        return out

    def _make_release(self, epoch: int, trainer: int):
        def release():
            if self._closed:
                return
            self.C.set_device(self.device_index)
            # Stream-ordered after everything the trainer enqueued on its stream.
            stream = self.torch.cuda.current_stream().cuda_stream
            targets = [self._consumed_ptr(r, trainer) for r in range(self.world)]
            self.C.signal_flags(targets, epoch + 1, stream)
            self.launches += 1
        return release

    def release_epoch(self, epoch, buffers):
        for b in buffers.values():
            b.release()

    # ------------------------------------------------------------------
    # helpers used by tests / bench
    # ------------------------------------------------------------------
    def check_error(self):
        err = self.poller.read(self.arena + self.off_error, 1)[0]
        if err:
            raise TimeoutError(f"device-side flag wait timed out (flag {err - 1})")

    def key_checksum(self, packed, field_name: str = "key") -> Tuple[int, int]:
        """Order-independent (sum, xor-hash) of an int64 field of packed rows."""
        torch = self.torch
        f = self.layout.field(field_name)
        out = torch.zeros(2, dtype=torch.int64, device=packed.device)
        self.C.key_checksum(packed.data_ptr(), packed.shape[0], self.layout.row_pitch,
                            f.offset, out.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        s, x = out.cpu().tolist()
        return s & (2**64 - 1), x & (2**64 - 1)

    def batch_sum(self, packed, field_name: str, out=None):
        """Device-side fp64 sum of an fp32 field of a packed batch (bench sink)."""
        torch = self.torch
        f = self.layout.field(field_name)
        if out is None:
            out = torch.zeros(1, dtype=torch.float64, device=packed.device)
        self.C.batch_sum_f32(packed.data_ptr(), packed.shape[0], self.layout.row_pitch,
                             f.offset, out.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        return out

    def batch_sum_all(self, packed, out, nbytes: Optional[int] = None):
        """fp64 sum of every fp32 word of a packed fp32 batch, accumulated into
        ``out`` (1-element float64 CUDA tensor) on the current stream. ``packed``
        is the batch's byte matrix, or any view that starts at the batch's first
        byte together with ``nbytes`` (e.g. the ``packed_features`` matrix)."""
        if nbytes is None:
            nbytes = packed.shape[0] * packed.shape[1] * packed.element_size()
        self.C.batch_sum_all_f32(packed.data_ptr(), nbytes, out.data_ptr(),
                                 self.torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        return out

    def bytes_in_use(self) -> int:
        if self._closed:
            return 0
        return int(self.arena_bytes + self.src_arena_bytes)

    def quiesce(self):
        """End of the last epoch: drain our streams, wait (collectively) until no
        peer can still push into us, and drop the peer mappings. Our own arena
        stays allocated so batches the user still holds remain valid until
        ``close()``."""
        if self._closed or getattr(self, "_quiesced", False):
            return
        self._quiesced = True
        C = self.C
        C.set_device(self.device_index)
        C.stream_synchronize(self.shuffle_stream)
        C.stream_synchronize(self.copy_stream)
        self.torch.cuda.synchronize(self.device_index)
        if self.world > 1:
            bootstrap.barrier(self.pg)
        for base in self._opened:
            try:
                C.ipc_close_handle(base)
            except Exception:
                pass
        self._opened = []

    def close(self):
        """Free every device / pinned allocation (not collective once
        ``quiesce`` ran)."""
        if self._closed:
            return
        try:
            self.quiesce()
        finally:
            self._closed = True
            stats_mod.unregister_bytes_used_source(self._bytes_fn)
            C = self.C
            for ev in list(self._events.values()) + list(self._first_pass_events.values()):
                C.event_destroy(ev[0])
                C.event_destroy(ev[1])
            self._events.clear()
            self._first_pass_events.clear()
            if self.resident != "hbm":
                for e in self.h2d_done + self.buf_free:
                    C.event_destroy(e)
            if hasattr(self, "_disk_pool"):
                self._disk_pool.shutdown(wait=True)
                for slot in self._disk_ring:
                    C.event_destroy(slot["copied"])
            for ptr in getattr(self, "_pinned", []):
                C.pinned_free(ptr)
            self._pinned = []
            C.pinned_free(self._desc_host_ptr)
            C.device_free(self.desc_arena)
            C.device_free(self.src_arena)
            if self._symm is None:
                C.device_free(self.arena)
            self._symm = None
            C.stream_destroy(self.shuffle_stream)
            C.stream_destroy(self.copy_stream)

    def __del__(self):
        try:
            if not self._closed and getattr(self, "_quiesced", False):
                self.close()
        except Exception:
            pass


def _pad(b: bytes, a: int = 256) -> bytes:
    return b + b"\0" * ((-len(b)) % a) if b else b"\0" * a
