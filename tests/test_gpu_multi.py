"""Multi-GPU tests (>= 2 B200): the P2P scatter over NVLink and the NCCL
baseline both reproduce the numpy golden on every rank. One process per GPU,
``torch.multiprocessing.spawn`` standing in for the cluster."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, files, ncols, dst_code, fp8, exchange, resident, out_dir,
            peer_alloc="symm", chunk_passes=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    torch.cuda.set_device(rank)
    dist.init_process_group("cuda:nccl,cpu:gloo", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    cols = [f"f{i}" for i in range(ncols - 1)] + ["labels"]

    def fn(schema):
        return L.build_layout([(c, schema[c][0], dst_code, 1) for c in cols],
                              fp8_block_scale=fp8)
    plan_args = dict(num_trainers=world, num_reducers=world * (chunk_passes or 1),
                     batch_size=1000, drop_last=False)
    gold = CpuShuffleEngine(files, plan_args, fn, 42)          # all trainers, one process
    opts = dict(exchange=exchange, resident=resident, peer_alloc=peer_alloc)
    if chunk_passes:
        opts["chunk_passes"] = chunk_passes
    if resident == "host":
        opts["stream_chunk_rows"] = 4096
    dev = DeviceShuffleEngine(files, plan_args, fn, 42, rank=rank, world=world, **opts)
    ok = True
    for epoch in range(4):                                      # > window: slots get reused
        gb = gold.start_epoch(epoch)
        db = dev.start_epoch(epoch)
        gb[rank].wait(120)
        if chunk_passes:
            # K7: every chunk is complete (all sources, over NVLink) as soon as ITS
            # pass's flags have fired - checked chunk by chunk, first to last
            for a, b in dev.plan.trainer_chunks(rank):
                db[rank].wait(120, row_stop=b)
                part = db[rank].data[a:b].cpu().numpy()
                ok = ok and np.array_equal(part, gb[rank].data[a:b])
        db[rank].wait(120)
        got = db[rank].data.cpu().numpy()
        ok = ok and np.array_equal(got, gb[rank].data)
        db[rank].release()
    torch.cuda.synchronize()
    dev.close()
    gold.close()
    with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def _files(tmp_path_factory, ncols, nrows=60_013):
    from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec
    d = tmp_path_factory.mktemp(f"multi{ncols}")
    files, _ = generate_data(nrows, 4, 2, 0.0, str(d), data_spec=float_spec(ncols, np.float32),
                             seed=ncols)
    return files


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs >= 2 CUDA devices")
@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange,resident,dst,fp8,peer_alloc,passes", [
    ("p2p", "hbm", 7, False, "symm", None), ("p2p", "host", 7, False, "symm", None),
    ("p2p", "hbm", 6, False, "symm", None), ("p2p", "hbm", 9, True, "symm", None),
    ("p2p", "hbm", 7, False, "ipc", None), ("nccl", "hbm", 7, False, "symm", None),
    ("p2p", "hbm", 7, False, "symm", 3)])
def test_multi_gpu_matches_golden(tmp_path_factory, tmp_path, exchange, resident, dst, fp8,
                                  peer_alloc, passes):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    files = _files(tmp_path_factory, 64)
    mp.spawn(_worker, args=(world, _free_port(), files, 64, dst, fp8, exchange, resident,
                            str(tmp_path), peer_alloc, passes), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok_{r}").read() == "1", f"rank {r} mismatch"


# ---------------------------------------------------------------------------
# the multi-rank protocol on ONE GPU (runs on single-GPU test boxes too)
# ---------------------------------------------------------------------------

def _worker_same_gpu(rank, world, port, files, ncols, out_dir, chunk_passes):
    """Two ranks = two processes sharing cuda:0: the peer mapping (legacy CUDA IPC works
    between processes on one device), the per-pass produced flags, the consumed-flag slot
    reuse gate and the exchange itself run exactly as on N GPUs - only the stores that would
    cross NVLink land in the same HBM. Process group: gloo (NCCL refuses two ranks per GPU)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK="0")
    import torch
    import torch.distributed as dist
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols = [f"f{i}" for i in range(ncols - 1)] + ["labels"]

    def fn(schema):
        return L.build_layout([(c, schema[c][0], L.DT_F32, 1) for c in cols])
    plan_args = dict(num_trainers=world, num_reducers=world * (chunk_passes or 1),
                     batch_size=1000, drop_last=False)
    gold = CpuShuffleEngine(files, plan_args, fn, 42)
    opts = dict(peer_alloc="ipc", device_index=0, numa_bind=False)
    if chunk_passes:
        opts["chunk_passes"] = chunk_passes
    dev = DeviceShuffleEngine(files, plan_args, fn, 42, rank=rank, world=world, **opts)
    ok = True
    for epoch in range(5):                                      # > window: slots get reused
        gb = gold.start_epoch(epoch)
        db = dev.start_epoch(epoch)
        gb[rank].wait(120)
        for a, b in dev.plan.trainer_chunks(rank):
            db[rank].wait(120, row_stop=b)
            ok = ok and np.array_equal(db[rank].data[a:b].cpu().numpy(), gb[rank].data[a:b])
        db[rank].wait(120)
        ok = ok and np.array_equal(db[rank].data.cpu().numpy(), gb[rank].data)
        db[rank].release()
    torch.cuda.synchronize()
    dev.check_error()
    dev.close()
    gold.close()
    with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("passes", [2])      # 2 passes: every per-pass flag path is exercised
def test_two_ranks_on_one_gpu_match_golden(tmp_path_factory, tmp_path, passes):
    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    files = _files(tmp_path_factory, 64, nrows=40_009)
    mp.spawn(_worker_same_gpu, args=(2, _free_port(), files, 64, str(tmp_path), passes),
             nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"ok_{r}").read() == "1", f"rank {r} mismatch"
