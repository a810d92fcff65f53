"""ShufflingDataset / TorchShufflingDataset / shuffle() on the CPU backend:
the property tests the reference lacks: exactly-once,
per-epoch variation, determinism, balance, drop_last, dtype/shape contract."""
import threading

import numpy as np
import pandas as pd
import pytest
import torch

from ray_shuffling_data_loader_b200 import (ShufflingDataset, TorchShufflingDataset,
                                            shuffle)
from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
from ray_shuffling_data_loader_b200.shuffle import BatchConsumer


def _epoch_keys(ds, epoch):
    ds.set_epoch(epoch)
    batches = list(ds)
    return batches, np.concatenate([b["key"].to_numpy() for b in batches])


def test_single_trainer_exactly_once(small_dataset):
    filenames, n = small_dataset
    ds = ShufflingDataset(filenames, 3, 1, 1000, 0, num_reducers=2, seed=11,
                          backend="cpu", queue_name="t1")
    seen = []
    for epoch in range(3):
        batches, keys = _epoch_keys(ds, epoch)
        assert isinstance(batches[0], pd.DataFrame)
        assert list(batches[0].columns) == ["key"] + list(DATA_SPEC.keys())
        assert [len(b) for b in batches[:-1]] == [1000] * (len(batches) - 1)
        assert len(batches[-1]) == n % 1000
        assert np.array_equal(np.sort(keys), np.arange(n))      # exactly once
        assert not np.array_equal(keys, np.arange(n))           # and shuffled
        seen.append(keys)
    assert not np.array_equal(seen[0], seen[1])                 # new permutation per epoch
    # rows stay intact: every column still belongs to its key
    full = pd.concat([pd.read_parquet(f) for f in filenames]).set_index("key")
    b = batches[0].set_index("key")
    assert np.array_equal(b["embeddings_name0"].to_numpy(),
                          full.loc[b.index, "embeddings_name0"].to_numpy())
    assert np.array_equal(b["labels"].to_numpy(), full.loc[b.index, "labels"].to_numpy())


def test_set_epoch_contract(small_dataset):
    filenames, n = small_dataset
    ds = ShufflingDataset(filenames, 2, 1, 4000, 0, num_reducers=3, seed=1,
                          backend="cpu", queue_name="t2")
    with pytest.raises(ValueError):
        next(iter(ds))
    ds.set_epoch(0)
    assert sum(len(b) for b in ds) == n
    with pytest.raises(ValueError):       # same epoch again
        next(iter(ds))
    ds.set_epoch(1)
    assert sum(len(b) for b in ds) == n


def test_determinism_and_drop_last(small_dataset):
    filenames, n = small_dataset
    runs = []
    for i in range(2):
        ds = ShufflingDataset(filenames, 1, 1, 1024, 0, num_reducers=4, seed=99,
                              drop_last=True, backend="cpu", queue_name=f"t3-{i}")
        batches, keys = _epoch_keys(ds, 0)
        assert all(len(b) == 1024 for b in batches)
        assert len(batches) == n // 1024
        runs.append(keys)
    assert np.array_equal(runs[0], runs[1])


@pytest.mark.parametrize("num_trainers,num_reducers", [(2, 2), (4, 6), (3, 1)])
def test_multi_trainer_disjoint_and_balanced(small_dataset, num_trainers, num_reducers):
    filenames, n = small_dataset
    name = f"t4-{num_trainers}-{num_reducers}"
    dss = [ShufflingDataset(filenames, 2, num_trainers, 500, r,
                            num_reducers=num_reducers, seed=5, backend="cpu",
                            queue_name=name) for r in range(num_trainers)]
    for epoch in range(2):
        per_rank = [None] * num_trainers

        def run(r):
            per_rank[r] = _epoch_keys(dss[r], epoch)[1]
        threads = [threading.Thread(target=run, args=(r,)) for r in range(num_trainers)]
        [t.start() for t in threads]
        [t.join(timeout=120) for t in threads]
        sizes = [len(k) for k in per_rank]
        assert max(sizes) - min(sizes) <= 1                      # balanced +-1 row
        assert np.array_equal(np.sort(np.concatenate(per_rank)), np.arange(n))


def test_state_dict_resume(small_dataset):
    filenames, n = small_dataset
    ds = ShufflingDataset(filenames, 3, 1, 700, 0, num_reducers=2, seed=21,
                          backend="cpu", queue_name="t5a")
    ds.set_epoch(0)
    list(ds)
    ds.set_epoch(1)
    it = iter(ds)
    first = [next(it)["key"].to_numpy() for _ in range(4)]
    state = ds.state_dict()
    assert state["epoch"] == 1 and state["batches_consumed"] == 4
    rest_ref = [b["key"].to_numpy() for b in it]
    ds.set_epoch(2)
    e2_ref = np.concatenate([b["key"].to_numpy() for b in ds])
    # resume in a fresh dataset
    ds2 = ShufflingDataset(filenames, 3, 1, 700, 0, num_reducers=2, seed=state["seed"],
                           backend="cpu", queue_name="t5b", start_epoch=state["epoch"])
    ds2.load_state_dict(state)
    ds2.set_epoch(1)
    rest = [b["key"].to_numpy() for b in ds2]
    assert len(rest) == len(rest_ref)
    assert all(np.array_equal(a, b) for a, b in zip(rest, rest_ref))
    ds2.set_epoch(2)
    assert np.array_equal(np.concatenate([b["key"].to_numpy() for b in ds2]), e2_ref)


def test_early_break_does_not_deadlock(small_dataset):
    filenames, n = small_dataset
    ds = ShufflingDataset(filenames, 3, 1, 1000, 0, num_reducers=2, seed=3,
                          backend="cpu", queue_name="t6", max_concurrent_epochs=1)
    ds.set_epoch(0)
    for i, _ in enumerate(ds):
        if i == 1:
            break
    ds.set_epoch(1)
    assert sum(len(b) for b in ds) == n
    ds.set_epoch(2)
    assert sum(len(b) for b in ds) == n


def test_torch_dataset_contract(small_dataset):
    filenames, n = small_dataset
    feature_columns = list(DATA_SPEC.keys())
    feature_types = [torch.int64 if np.dtype(d).kind == "i" else torch.float64
                     for _, _, d in DATA_SPEC.values()]
    label_column = feature_columns.pop()
    label_type = feature_types.pop()
    ds = TorchShufflingDataset(
        filenames, 2, 1, 2048, 0, num_reducers=2, feature_columns=feature_columns,
        feature_types=feature_types, label_column=label_column, label_type=label_type,
        seed=17, backend="cpu", queue_name="t7")
    full = pd.concat([pd.read_parquet(f) for f in filenames])
    for epoch in range(2):
        ds.set_epoch(epoch)
        total = 0
        for features, label in ds:
            assert isinstance(features, list) and len(features) == 19
            b = label.shape[0]
            assert label.shape == (b, 1) and label.dtype == torch.float64
            for t in features:
                assert t.shape == (b, 1) and t.dtype == torch.int64
            total += b
        assert total == n
    # values: (embeddings_name12, labels) pairs are (almost surely) unique row ids
    pairs = set(zip(full["embeddings_name12"].tolist(), full["labels"].tolist()))
    got = set(zip(features[12][:, 0].tolist(), label[:, 0].tolist()))
    assert got <= pairs


def test_torch_dataset_defaults_and_casts(small_dataset):
    filenames, n = small_dataset
    ds = TorchShufflingDataset(filenames, 1, 1, 5000, 0, num_reducers=1,
                               feature_columns=["embeddings_name0", "one_hot0"],
                               label_column="labels", seed=2, backend="cpu",
                               queue_name="t8")
    ds.set_epoch(0)
    features, label = next(iter(ds))
    assert [t.dtype for t in features] == [torch.float32, torch.float32]  # default torch.float
    assert label.dtype == torch.float32 and label.shape == (5000, 1)
    assert float(features[1].max()) <= 2.0 and float(features[0].max()) <= 2384.0


def test_packed_features_and_bf16(float_dataset):
    filenames, n = float_dataset
    cols = [f"f{i}" for i in range(15)]
    ds = TorchShufflingDataset(filenames, 1, 1, 1000, 0, num_reducers=2,
                               feature_columns=cols, feature_types=[torch.bfloat16] * 15,
                               label_column="labels", label_type=torch.bfloat16,
                               packed_features=True, seed=4, backend="cpu",
                               queue_name="t9")
    ds.set_epoch(0)
    full = pd.concat([pd.read_parquet(f) for f in filenames])
    ref = {round(float(torch.tensor(v, dtype=torch.float32).to(torch.bfloat16)), 6)
           for v in full["f3"].to_numpy()[:50]}
    rows = 0
    seen = set()
    for feats, label in ds:
        assert feats.dtype == torch.bfloat16 and feats.shape[1] == 15
        assert label.shape == (feats.shape[0], 1)
        rows += feats.shape[0]
        seen |= {round(float(v), 6) for v in feats[:, 3].float()}
    assert rows == n
    assert ref <= seen


def test_list_column_feature_shapes(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    n = 300
    emb = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
    files = []
    for i, sl in enumerate([slice(0, 100), slice(100, 300)]):
        tbl = pa.table({
            "key": pa.array(np.arange(n)[sl]),
            "emb": pa.FixedSizeListArray.from_arrays(pa.array(emb[sl].reshape(-1)), 4),
            "y": pa.array(np.arange(n, dtype=np.float64)[sl]),
        })
        fn = str(tmp_path / f"l{i}.parquet")
        pq.write_table(tbl, fn)
        files.append(fn)
    ds = TorchShufflingDataset(files, 1, 1, 128, 0, num_reducers=2,
                               feature_columns=["emb"], feature_shapes=[(2, 2)],
                               label_column="y", seed=0, backend="cpu", queue_name="t10")
    ds.set_epoch(0)
    ys = []
    for (e,), y in ds:
        assert e.shape[1:] == (2, 2) and y.shape[1] == 1
        # row integrity: emb row == 4*y + [0,1,2,3]
        assert torch.equal(e.reshape(-1, 4)[:, 0], y[:, 0] * 4)
        ys.append(y[:, 0])
    assert sorted(torch.cat(ys).tolist()) == list(range(n))
    # plain dataset yields object cells of ndarrays for list columns
    ds2 = ShufflingDataset(files, 1, 1, 300, 0, num_reducers=1, seed=0, backend="cpu",
                           queue_name="t10b")
    ds2.set_epoch(0)
    df = next(iter(ds2))
    assert isinstance(df["emb"].iloc[0], np.ndarray) and df["emb"].iloc[0].shape == (4,)


class _RecordingConsumer(BatchConsumer):
    def __init__(self, num_trainers, window, stats=None):
        self.rows = {}
        self.window = window
        self.done = {}
        self.stats = stats
        self.lock = threading.Lock()
        self.order = []

    def consume(self, rank, epoch, batches):
        for chunk in batches:
            df = chunk.to_pandas()
            with self.lock:
                self.rows.setdefault((epoch, rank), []).append(df["key"].to_numpy())
            if self.stats is not None:
                self.stats.consume_batch(epoch, len(chunk))

    def producer_done(self, rank, epoch):
        self.done[(epoch, rank)] = True
        if self.stats is not None:
            self.stats.consume_done(epoch)

    def wait_until_ready(self, epoch):
        self.order.append(epoch)

    def wait_until_all_epochs_done(self):
        pass


def test_shuffle_api_with_custom_consumer_and_stats(small_dataset, tmp_path):
    from ray_shuffling_data_loader_b200.stats import (TrialStatsCollector,
                                                      ObjectStoreStatsCollector,
                                                      process_stats)
    filenames, n = small_dataset
    num_epochs, num_reducers, num_trainers = 3, 5, 2
    stats = TrialStatsCollector(num_epochs, len(filenames), num_reducers, num_trainers)
    consumer = _RecordingConsumer(num_trainers, 2, stats)
    with ObjectStoreStatsCollector(0.05) as store:
        duration = shuffle(filenames, consumer, num_epochs, num_reducers, num_trainers,
                           stats, seed=8, backend="cpu")
    assert duration > 0 and consumer.order == [0, 1, 2]
    for epoch in range(num_epochs):
        keys = np.concatenate([np.concatenate(consumer.rows[(epoch, r)])
                               for r in range(num_trainers)])
        assert np.array_equal(np.sort(keys), np.arange(n))
        assert sum(len(consumer.rows[(epoch, r)]) for r in range(num_trainers)) == num_reducers
    trial = stats.get_stats(timeout=10)
    assert len(trial.epoch_stats) == num_epochs and trial.duration == duration
    assert len(trial.epoch_stats[0].reduce_stats.task_durations) == num_reducers
    assert len(trial.epoch_stats[0].consume_stats.time_to_consumes) == num_reducers
    process_stats([(trial, store.get_stats())], True, str(tmp_path), False, False, False,
                  n, len(filenames), 2, 1000, num_reducers, num_trainers, num_epochs, 2)
    from ray_shuffling_data_loader_b200.stats import human_readable_big_num as hr
    tag = f"{hr(n)}_rows_{hr(1000)}_batch_size.csv"
    assert tag == "10.0K_rows_1K_batch_size.csv"
    trial_csv = pd.read_csv(tmp_path / f"trial_stats_{tag}")
    epoch_csv = pd.read_csv(tmp_path / f"epoch_stats_{tag}")
    cons_csv = pd.read_csv(tmp_path / f"consumer_stats_{tag}")
    assert len(trial_csv.columns) == 45 and len(trial_csv) == 1   # reference schema
    assert len(epoch_csv.columns) == 31 and len(epoch_csv) == num_epochs
    assert len(cons_csv.columns) == 10
    assert abs(trial_csv["row_throughput"][0] - num_epochs * n / duration) < 1e-6


# ---------------------------------------------------------------------------
# round-2 additions
# ---------------------------------------------------------------------------

def test_pandas_written_list_column(tmp_path):
    """Parquet written by pandas from ndarray-valued cells is a plain
    ``list<item>`` column (not fixed_size_list): its width comes from the footer at
    scan time, so layouts built before any decode are correct."""
    n = 240
    img = np.arange(n * 6, dtype=np.float32).reshape(n, 6)
    files = []
    for i, sl in enumerate([slice(0, 100), slice(100, 240)]):
        df = pd.DataFrame({"key": np.arange(n)[sl], "img": list(img[sl]),
                           "y": np.arange(n, dtype=np.float64)[sl]})
        fn = str(tmp_path / f"p{i}.parquet")
        df.to_parquet(fn)
        files.append(fn)
    from ray_shuffling_data_loader_b200.runtime import ingest
    assert ingest.scan_files(files).schema["img"][1] == 6
    for native in (True, False):
        ds = TorchShufflingDataset(files, 1, 1, 64, 0, num_reducers=2, feature_columns=["img"],
                                   feature_shapes=[(2, 3)], label_column="y", seed=3,
                                   backend="cpu", queue_name=f"plist{native}", native=native)
        ds.set_epoch(0)
        ys = []
        for (e,), y in ds:
            assert e.shape[1:] == (2, 3)
            assert torch.equal(e.reshape(-1, 6)[:, 0], y[:, 0] * 6)
            ys.append(y[:, 0])
        assert sorted(torch.cat(ys).tolist()) == list(range(n))
    # plain dataset: DataFrame with ndarray cells, like pandas reading the file
    ds2 = ShufflingDataset(files, 1, 1, 240, 0, num_reducers=1, seed=0, backend="cpu",
                           queue_name="plist-df")
    ds2.set_epoch(0)
    df = next(iter(ds2))
    assert df["img"].iloc[0].shape == (6,)


def test_ragged_list_column_is_rejected(tmp_path):
    df = pd.DataFrame({"key": np.arange(4), "v": [np.zeros(2, np.float32), np.zeros(3, np.float32),
                                                  np.zeros(2, np.float32), np.zeros(1, np.float32)]})
    fn = str(tmp_path / "ragged.parquet")
    df.to_parquet(fn)
    ds = ShufflingDataset([fn], 1, 1, 4, 0, num_reducers=1, seed=0, backend="cpu",
                          queue_name="ragged")
    ds.set_epoch(0)
    with pytest.raises(Exception, match="constant length|same length|shuffle driver"):
        list(ds)


def _torch_rank1_child(qdir, files, out_path):
    import os
    os.environ["RSDL_B200_QUEUE_DIR"] = qdir
    import torch
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    ds = TorchShufflingDataset(files, 2, 2, 500, 1, num_reducers=4,
                               feature_columns=["key", "embeddings_name0"],
                               feature_types=[torch.int64, torch.float32],
                               label_column="labels", backend="cpu", queue_name="torch-2rank")
    keys = []
    for epoch in range(2):
        ds.set_epoch(epoch)
        for (k, e), y in ds:
            assert k.dtype == torch.int64 and e.shape[1] == 1 and y.shape[1] == 1
            keys.append(k[:, 0].clone())
    torch.save(torch.cat(keys), out_path)


def test_torch_dataset_connecting_rank_other_process(small_dataset, tmp_path, monkeypatch):
    """The reference's canonical launch: rank 0 owns queue + shuffle, rank 1 (another
    process, no torch.distributed) connects by name and gets tensors (used
    to raise 'needs the layout of the owning process' and hang rank 0)."""
    import multiprocessing as mp
    files, n = small_dataset
    monkeypatch.setenv("RSDL_B200_QUEUE_DIR", str(tmp_path))
    out = str(tmp_path / "rank1.pt")
    ds = TorchShufflingDataset(files, 2, 2, 500, 0, num_reducers=4,
                               feature_columns=["key", "embeddings_name0"],
                               feature_types=[torch.int64, torch.float32],
                               label_column="labels", seed=5, backend="cpu",
                               queue_name="torch-2rank")
    p = mp.get_context("spawn").Process(target=_torch_rank1_child,
                                        args=(str(tmp_path), files, out))
    p.start()
    mine = []
    for epoch in range(2):
        ds.set_epoch(epoch)
        for (k, e), y in ds:
            mine.append(k[:, 0].clone())
    p.join(timeout=120)
    assert p.exitcode == 0
    theirs = torch.load(out)
    both = torch.cat(mine + [theirs]).tolist()
    assert sorted(both) == sorted(list(range(n)) * 2)


def test_tensor_spec_and_host_converter():
    from ray_shuffling_data_loader_b200.torch_dataset import (TensorSpec, convert_to_tensor,
                                                              dataframe_to_tensor_factory)
    spec = TensorSpec.build(["a", "img"], [None, (2, 2)], None, "y", None, torch.int64)
    assert spec.features[0].dtype == torch.float and spec.features[1].shape == (2, 2)
    assert spec.label.dtype == torch.int64 and spec.label.trailing() == (1,)
    assert TensorSpec.build("a", None, None, "y").features[0].name == "a"     # scalar promotion
    with pytest.raises(ValueError):
        TensorSpec.build(["a", "b"], [None], None, "y")
    with pytest.raises(TypeError):
        TensorSpec.build(["a"], None, [np.float32], "y")
    df = pd.DataFrame({"a": np.arange(5, dtype=np.int64),
                       "img": [np.full(4, i, dtype=np.float32) for i in range(5)],
                       "t": [(i, i) for i in range(5)],
                       "y": np.arange(5, dtype=np.float64)})
    feats, label = convert_to_tensor(df, spec)
    assert feats[0].shape == (5, 1) and feats[0].dtype == torch.float32
    assert feats[1].shape == (5, 2, 2) and float(feats[1][3, 1, 1]) == 3.0
    assert label.dtype == torch.int64 and label.shape == (5, 1)
    conv = dataframe_to_tensor_factory(["t"], None, [torch.int32], "y", 1)
    (t,), y = conv(df)
    assert t.shape == (5, 2) and t.dtype == torch.int32 and y.shape == (5, 1)
    with pytest.raises(TypeError):
        convert_to_tensor(pd.DataFrame({"s": ["x", "y"], "y": [0.0, 1.0]}),
                          feature_columns=["s"], label_column="y")
