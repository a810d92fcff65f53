"""Multi-process (gloo, world_size=2) run of the full stack on CPU: the
host-side logic of distributed mode, exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, filenames, n, qdir, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      RSDL_B200_QUEUE_DIR=qdir)
    import torch
    import torch.distributed as dist
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = TorchShufflingDataset(
        filenames, 2, world, 512, rank, num_reducers=4,
        feature_columns=["key", "embeddings_name0"],
        feature_types=[torch.int64, torch.float32],
        label_column="labels", seed=None, backend="cpu")
    for epoch in range(2):
        ds.set_epoch(epoch)
        keys = []
        for (key, emb), label in ds:
            assert key.shape[1] == 1 and emb.dtype == torch.float32
            keys.append(key[:, 0].clone())
        np.save(os.path.join(out_dir, f"keys_{epoch}_{rank}.npy"), torch.cat(keys).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_two_ranks_exactly_once(small_dataset, tmp_path):
    filenames, n = small_dataset
    world = 2
    port = _free_port()
    qdir = str(tmp_path / "q")
    os.makedirs(qdir)
    mp.spawn(_worker, args=(world, port, filenames, n, qdir, str(tmp_path)),
             nprocs=world, join=True)
    for epoch in range(2):
        parts = [np.load(tmp_path / f"keys_{epoch}_{r}.npy") for r in range(world)]
        assert abs(len(parts[0]) - len(parts[1])) <= 1
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
    e0 = np.concatenate([np.load(tmp_path / f"keys_0_{r}.npy") for r in range(world)])
    e1 = np.concatenate([np.load(tmp_path / f"keys_1_{r}.npy") for r in range(world)])
    assert not np.array_equal(e0, e1)
