"""GPU tests of the public API: datasets yield CUDA batches with the reference's
contract; exactly-once is proven on device with key_checksum (K13)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

from ray_shuffling_data_loader_b200 import ShufflingDataset, TorchShufflingDataset
from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
from ray_shuffling_data_loader_b200.runtime.chunks import DeviceBatch


def test_device_batches_exactly_once(small_dataset):
    files, n = small_dataset
    ds = ShufflingDataset(files, 4, 1, 1000, 0, num_reducers=3, seed=31, queue_name="g1")
    assert ds.engine.device == "cuda"
    want_sum = n * (n - 1) // 2
    prev = None
    for epoch in range(4):
        ds.set_epoch(epoch)
        keys, total_sum, total_xor = [], 0, 0
        for b in ds:
            assert isinstance(b, DeviceBatch) and b.packed.is_cuda
            assert len(b) == 1000 or len(b) == n % 1000
            s, x = ds.engine.key_checksum(b.packed)
            total_sum = (total_sum + s) % 2**64
            total_xor ^= x
            keys.append(b["key"].clone())
        keys = torch.cat(keys).cpu().numpy()
        assert total_sum == want_sum
        assert np.array_equal(np.sort(keys), np.arange(n))
        assert prev is None or not np.array_equal(keys, prev)
        prev = keys
    # batches held after the final epoch stay valid until the dataset is closed
    df = b.to_pandas()
    assert list(df.columns) == ["key"] + list(DATA_SPEC.keys())
    assert np.array_equal(df["key"].to_numpy(), keys[-len(df):])
    ds.close()


def test_gpu_equals_cpu_batches(small_dataset):
    files, n = small_dataset
    g = ShufflingDataset(files, 2, 1, 777, 0, num_reducers=2, seed=99, queue_name="g2a")
    c = ShufflingDataset(files, 2, 1, 777, 0, num_reducers=2, seed=99, backend="cpu",
                         queue_name="g2b")
    for epoch in range(2):
        g.set_epoch(epoch); c.set_epoch(epoch)
        for gb, cb in zip(g, c):
            assert np.array_equal(gb["key"].cpu().numpy(), cb["key"].to_numpy())
            assert np.array_equal(gb["labels"].cpu().numpy(), cb["labels"].to_numpy())


def test_torch_dataset_cuda_contract(float_dataset):
    files, n = float_dataset
    cols = [f"f{i}" for i in range(15)]
    ds = TorchShufflingDataset(files, 3, 1, 500, 0, num_reducers=2, feature_columns=cols,
                               label_column="labels", seed=3, queue_name="g3",
                               max_concurrent_epochs=2)
    import pandas as pd
    full = pd.concat([pd.read_parquet(f) for f in files])
    for epoch in range(3):
        ds.set_epoch(epoch)
        rows, lab_sum = 0, 0.0
        for feats, label in ds:
            assert len(feats) == 15 and all(t.is_cuda and t.dtype == torch.float32 for t in feats)
            assert all(t.shape == (label.shape[0], 1) for t in feats)
            rows += label.shape[0]
            lab_sum += float(label.double().sum())
        assert rows == n
        assert abs(lab_sum - float(full["labels"].astype(np.float64).sum())) < 1e-3


def test_torch_dataset_packed_bf16_and_fp8(float_dataset):
    files, n = float_dataset
    cols = [f"f{i}" for i in range(15)]
    ds = TorchShufflingDataset(files, 1, 1, 1024, 0, num_reducers=2, feature_columns=cols,
                               feature_types=[torch.bfloat16] * 15, label_column="labels",
                               label_type=torch.float32, packed_features=True, seed=3,
                               queue_name="g4")
    ds.set_epoch(0)
    total = 0
    for feats, label in ds:
        assert feats.dtype == torch.bfloat16 and feats.shape[1] == 15 and feats.is_cuda
        assert label.dtype == torch.float32
        total += feats.shape[0]
    assert total == n
    if hasattr(torch, "float8_e4m3fn"):
        ds = TorchShufflingDataset(files, 1, 1, 1024, 0, num_reducers=2, feature_columns=cols,
                                   feature_types=[torch.float8_e4m3fn] * 15,
                                   label_column="labels", label_type=torch.float32,
                                   packed_features=True, fp8_block_scale=True, seed=3,
                                   queue_name="g5")
        ds.set_epoch(0)
        for (payload, scales), label in ds:
            assert payload.dtype == torch.float8_e4m3fn and payload.shape[1] == 15
            assert scales.dtype == torch.uint8 and scales.shape[1] == 1
            deq = payload.float() * torch.exp2(scales.float() - 127)
            assert float(deq.max()) <= 1.0 + 1e-6 and float(deq.min()) >= 0.0


def test_window_backpressure_many_epochs(float_dataset):
    """More epochs than ring slots: slots are reused only after release."""
    files, n = float_dataset
    ds = ShufflingDataset(files, 7, 2, 600, 0, num_reducers=4, seed=8, queue_name="g6",
                          max_concurrent_epochs=2)
    ds1 = ShufflingDataset(files, 7, 2, 600, 1, num_reducers=4, seed=8, queue_name="g6")
    import threading
    out = {}

    def run(d, r):
        try:
            sums = []
            for epoch in range(7):
                d.set_epoch(epoch)
                ks = [b["key"].clone() for b in d]
                sums.append(torch.cat(ks).cpu().numpy())
            out[r] = sums
        except BaseException as e:   # never leave the other trainer blocked
            out[r] = e
            ds.close()
    t = threading.Thread(target=run, args=(ds1, 1))
    t.start()
    run(ds, 0)
    t.join(timeout=120)
    for r in (0, 1):
        assert not isinstance(out.get(r), BaseException), out[r]
    for epoch in range(7):
        keys = np.concatenate([out[0][epoch], out[1][epoch]])
        assert np.array_equal(np.sort(keys), np.arange(n)), epoch


def test_resume_with_start_epoch_beyond_window(float_dataset):
    """a fresh process resuming at start_epoch >= max_concurrent_epochs
    must not wait on consumed flags nobody will ever write (it used to sit in the
    slot-reuse gate for flag_timeout_s). Resumed batches equal the uninterrupted run's."""
    import time
    files, n = float_dataset
    cols = [f"f{i}" for i in range(15)]
    kw = dict(num_reducers=2, feature_columns=cols, label_column="labels", seed=77,
              packed_features=True, max_concurrent_epochs=2)
    full = TorchShufflingDataset(files, 6, 1, 700, 0, queue_name="res-a", **kw)
    want = {}
    for epoch in range(6):
        full.set_epoch(epoch)
        want[epoch] = [f.clone() for f, _ in full]
    state = None
    full.dataset.close()
    for mode in ("stream", "host"):
        t0 = time.monotonic()
        ds = TorchShufflingDataset(files, 6, 1, 700, 0, queue_name=f"res-{mode}", start_epoch=3,
                                   backpressure=mode, flag_timeout_s=20.0, **kw)
        ds.load_state_dict({"seed": 77, "epoch": 3, "batches_consumed": 2, "batch_size": 700})
        for epoch in range(3, 6):
            ds.set_epoch(epoch)
            got = [f.clone() for f, _ in ds]
            ref = want[epoch][2:] if epoch == 3 else want[epoch]
            assert len(got) == len(ref)
            assert all(torch.equal(a, b) for a, b in zip(got, ref)), (mode, epoch)
        ds.dataset.engine.check_error()
        assert time.monotonic() - t0 < 15.0, "resume sat in the back-pressure gate"
        ds.dataset.close()


def test_pandas_list_column_cuda(tmp_path):
    """plain list<item> Parquet column (pandas ndarray cells) through the GPU path."""
    import pandas as pd
    n = 3000
    img = np.arange(n * 32, dtype=np.float32).reshape(n, 32)
    fn = str(tmp_path / "pl.parquet")
    pd.DataFrame({"img": list(img), "y": np.arange(n, dtype=np.float64)}).to_parquet(fn)
    ds = TorchShufflingDataset([fn], 1, 1, 512, 0, num_reducers=2, feature_columns=["img"],
                               feature_shapes=[(4, 8)], label_column="y", seed=1,
                               queue_name="plist-gpu")
    ds.set_epoch(0)
    ys = []
    for (e,), y in ds:
        assert e.is_cuda and e.shape[1:] == (4, 8)
        assert torch.equal(e.reshape(-1, 32)[:, 0], y[:, 0] * 32)
        ys.append(y[:, 0])
    assert sorted(torch.cat(ys).tolist()) == list(range(n))
    ds.dataset.close()


def test_pandas_output_from_device_buffers(small_dataset):
    """output="pandas" on the GPU backend: one D2H copy and one native unpack per reducer
    chunk, batches are slices of the chunk frames - same DataFrames as the CPU backend."""
    import pandas as pd
    files, n = small_dataset
    g = ShufflingDataset(files, 2, 1, 777, 0, num_reducers=3, seed=5, output="pandas",
                         queue_name="gpd-a")
    c = ShufflingDataset(files, 2, 1, 777, 0, num_reducers=3, seed=5, backend="cpu",
                         queue_name="gpd-b")
    assert g.engine.device == "cuda"
    for epoch in range(2):
        g.set_epoch(epoch); c.set_epoch(epoch)
        rows = 0
        for gb, cb in zip(g, c):
            assert isinstance(gb, pd.DataFrame) and list(gb.columns) == list(cb.columns)
            pd.testing.assert_frame_equal(gb.reset_index(drop=True), cb.reset_index(drop=True))
            rows += len(gb)
        assert rows == n
    g.close(); c.close()
