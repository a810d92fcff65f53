"""Process-group bootstrap: one process per GPU, ``torch.distributed`` plumbing.

The reference's control plane is Ray (named actors, GCS, gRPC; reference
``batch_queue.py:63-65,358-380``, ``dataset.py:52-84``) and its example
launches workers with Horovod's ``RayExecutor``
(``examples/horovod/ray_torch_shuffle.py:336-345``). Here ranks are plain
processes started by ``torchrun`` / ``mp.spawn``; rendezvous, small-object
exchange (seeds, CUDA-IPC handles) and CPU-mode row exchange ride
``torch.distributed`` - NCCL over NVLink on GPUs, gloo on the CPU backend.
"""
from __future__ import annotations

import os
import random
from dataclasses import dataclass
from typing import Any, List, Optional


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    backend: Optional[str] = None

    @property
    def is_distributed(self) -> bool:
        return self.world > 1


def current_context() -> DistContext:
    """Describe the ambient ``torch.distributed`` state (no side effects)."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return DistContext()
    if dist.is_available() and dist.is_initialized():
        return DistContext(dist.get_rank(), dist.get_world_size(),
                           int(os.environ.get("LOCAL_RANK", dist.get_rank())),
                           dist.get_backend())
    return DistContext()


def init_from_env(backend: Optional[str] = None, device_id: Optional[int] = None,
                  timeout_s: float = 600.0) -> DistContext:
    """Initialise the default process group from torchrun-style env vars
    (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT) if needed."""
    import datetime
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return current_context()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world <= 1:
        return DistContext(0, 1, local_rank, None)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    use_cuda = torch.cuda.is_available()
    if backend is None:
        # gloo rides along for CPU objects even on GPU boxes.
        backend = "cuda:nccl,cpu:gloo" if use_cuda else "gloo"
    kwargs = {}
    if use_cuda:
        dev = local_rank if device_id is None else device_id
        torch.cuda.set_device(dev)
        kwargs["device_id"] = torch.device("cuda", dev)
    dist.init_process_group(backend=backend, rank=rank, world_size=world,
                            timeout=datetime.timedelta(seconds=timeout_s), **kwargs)
    return DistContext(rank, world, local_rank, backend)


def broadcast_object(obj: Any, src: int = 0, group=None) -> Any:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def all_gather_object(obj: Any, group=None) -> List[Any]:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def barrier(group=None) -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


def agree_on_seed(seed: Optional[int], group=None) -> int:
    """All ranks must evaluate the same permutation: rank 0's seed wins.
    ``None`` draws a fresh random seed (the reference is unseeded)."""
    if seed is None:
        seed = random.SystemRandom().getrandbits(63)
    return int(broadcast_object(int(seed), 0, group))
