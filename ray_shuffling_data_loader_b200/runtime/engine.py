"""Engine selection: CUDA (sm_100a kernels) when a GPU is visible, numpy otherwise."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.parallel import bootstrap


def cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def resolve_backend(backend: Optional[str]) -> str:
    if backend in (None, "auto"):
        return "cuda" if cuda_available() else "cpu"
    if backend not in ("cpu", "cuda"):
        raise ValueError(f"unknown backend {backend!r}")
    if backend == "cuda" and not cuda_available():
        raise RuntimeError("backend='cuda' requested but no CUDA device is visible")
    return backend


def make_engine(filenames: Sequence[str], *, num_trainers: int, num_reducers: int,
                batch_size: int, drop_last: bool = False,
                layout_fn: Optional[Callable] = None, seed: Optional[int] = None,
                backend: Optional[str] = None, stats_collector=None,
                rank: Optional[int] = None, world: Optional[int] = None,
                **options):
    """Build the shuffle engine for this process.

    Distributed mode (``torch.distributed`` initialised with
    ``world_size == num_trainers``): every rank owns 1/world of the rows and
    serves exactly its own trainer. Otherwise this single process serves all
    ``num_trainers`` logical trainers."""
    ctx = bootstrap.current_context()
    if world is None:
        world = ctx.world if ctx.world == num_trainers else 1
    if rank is None:
        rank = ctx.rank if world > 1 else 0
    if world > 1:
        seed = bootstrap.agree_on_seed(seed)     # collective: rank 0's seed wins
    elif seed is None:
        import random
        seed = random.SystemRandom().getrandbits(63)
    layout_fn = layout_fn or L.dataframe_layout
    plan_args = dict(num_trainers=num_trainers, num_reducers=num_reducers,
                     batch_size=batch_size, drop_last=drop_last)
    backend = resolve_backend(backend)
    recycle = bool(options.pop("recycle_buffers", False))       # host engine only
    if world > 1:
        # Collectives the engine issues from its *background* threads (the CPU
        # engine's gloo all_to_all on the "cpu-shuffle" thread, the NCCL baseline's
        # exchange on the shuffle-driver thread) must never share a communicator
        # with the training loop's DDP / all_reduce on the main thread: their
        # relative order would differ from rank to rank. Give them a private group
        # (collective: every rank builds its dataset, like agree_on_seed above).
        import torch.distributed as dist
        if backend == "cpu" and options.get("process_group") is None:
            options["process_group"] = dist.new_group(backend="gloo")
        elif backend == "cuda" and options.get("exchange") == "nccl" \
                and options.get("exchange_group") is None:
            options["exchange_group"] = dist.new_group(backend="nccl")
    if backend == "cuda":
        from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
        return DeviceShuffleEngine(filenames, plan_args, layout_fn, seed, rank=rank,
                                   world=world, stats_collector=stats_collector,
                                   **options)
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    cpu_opts = {k: v for k, v in options.items()
                if k in ("num_threads", "process_group", "index", "native")}
    return CpuShuffleEngine(filenames, plan_args, layout_fn, seed, rank=rank,
                            world=world, stats_collector=stats_collector,
                            recycle_buffers=recycle, **cpu_opts)
