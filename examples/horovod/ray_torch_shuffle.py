#!/usr/bin/env python
"""Self-launching trainer example with the reference's command line (C17).

The reference's ``examples/horovod/ray_torch_shuffle.py`` is started as one
script and brings up its own workers through Horovod's ``RayExecutor``
(``:336-345``), one GPU each. This launcher keeps that user experience on one
8xB200 host: it accepts the reference's flags (``:39-121``), spawns
``--num-workers`` (or ``--num-hosts x --num-workers-per-host``) processes with
``torch.multiprocessing``, one per GPU, rendezvous on 127.0.0.1, and runs
``examples/ddp/torch_shuffle.py::train_main`` in each - NCCL DDP instead of
Horovod, ``TorchShufflingDataset`` batches already on the device instead of
``.cuda()`` copies.

    python examples/horovod/ray_torch_shuffle.py --num-workers 8 --epochs 3

Flags that only made sense for Ray / Horovod are accepted and reported as
ignored (``--address``, ``--cpus-per-worker``, ``--use-adasum``,
``--gradient-predivide-factor``); ``--fp16-allreduce`` maps to DDP's bf16
gradient-compression hook.
"""
from __future__ import annotations

import argparse
import importlib.util
import os
import socket
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DDP_EXAMPLE = os.path.join(os.path.dirname(HERE), "ddp", "torch_shuffle.py")


def _load_ddp_example():
    spec = importlib.util.spec_from_file_location("rsdl_ddp_example", DDP_EXAMPLE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_parser(ddp) -> argparse.ArgumentParser:
    """The DDP example's parser plus the reference-only flags."""
    parser = ddp.parser
    parser.description = "Shuffling data loader example (reference CLI, self-launching)"
    parser.add_argument("--test-batch-size", type=int, default=250000)
    parser.add_argument("--fp16-allreduce", action="store_true", default=False,
                        help="compress gradients for the allreduce (bf16 hook under DDP)")
    parser.add_argument("--use-adasum", action="store_true", default=False)
    parser.add_argument("--gradient-predivide-factor", type=float, default=1.0)
    parser.add_argument("--num-workers", type=int, default=None)
    parser.add_argument("--num-hosts", type=int, default=None)
    parser.add_argument("--num-workers-per-host", type=int, default=None)
    parser.add_argument("--cpus-per-worker", type=int, default=1)
    parser.add_argument("--address", type=str, default="auto")
    return parser


def resolve_num_workers(args) -> int:
    """Same precedence as the reference (``:323-335``): an explicit worker count
    wins, otherwise hosts x workers-per-host; default: every visible GPU."""
    if args.num_workers is not None:
        if args.num_hosts is not None or args.num_workers_per_host is not None:
            raise ValueError("use either --num-workers or --num-hosts with --num-workers-per-host")
        return args.num_workers
    if args.num_hosts is not None or args.num_workers_per_host is not None:
        if args.num_hosts is None or args.num_workers_per_host is None:
            raise ValueError("--num-hosts and --num-workers-per-host go together")
        if args.num_hosts != 1:
            raise ValueError("this launcher drives one host (8xB200); use torchrun "
                             "--nnodes for more")
        return args.num_workers_per_host
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return max(1, n)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(local_rank: int, world: int, port: int, argv):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(local_rank),
                      LOCAL_RANK=str(local_rank), WORLD_SIZE=str(world))
    ddp = _load_ddp_example()
    args = build_parser(ddp).parse_args(argv)
    if args.fp16_allreduce:
        os.environ["RSDL_EXAMPLE_GRAD_COMPRESSION"] = "bf16"
    ddp.train_main(args)
    print(f"Done consuming batches on worker {local_rank}.")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ddp = _load_ddp_example()
    args = build_parser(ddp).parse_args(argv)
    world = resolve_num_workers(args)
    ignored = [f for f, on in (("--address", args.address != "auto"),
                               ("--cpus-per-worker", args.cpus_per_worker != 1),
                               ("--use-adasum", args.use_adasum),
                               ("--gradient-predivide-factor",
                                args.gradient_predivide_factor != 1.0)) if on]
    if ignored:
        print(f"note: {', '.join(ignored)} only apply to Ray/Horovod and are ignored here")
    if world == 1:
        _worker(0, 1, _free_port(), argv)
        return
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), argv), nprocs=world, join=True)


if __name__ == "__main__":
    main()
