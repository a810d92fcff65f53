"""The C++ host runtime (csrc/bindings.cpp) against the numpy golden: the
native CPU shuffle must reproduce the pure-numpy engine byte for byte, because
both share the definition of the bijection (perm.cuh == ops/perm.py)."""
import numpy as np
import pytest

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.ops import perm
from ray_shuffling_data_loader_b200.ops.plan import ShufflePlan

_C = pytest.importorskip("ray_shuffling_data_loader_b200._C")


@pytest.mark.parametrize("n,T,off,cnt", [(1000, 1, 0, 1000), (100_003, 8, 12_345, 40_000),
                                         (7, 3, 0, 7), (1 << 20, 5, 999, 70_000), (1, 1, 0, 1)])
def test_host_perm_positions_match_numpy(n, T, off, cnt):
    pool = _C.HostPool(3)
    for epoch in (0, 3):
        key = perm.make_key(n, 77, epoch)
        plan = ShufflePlan(n, T, T, 10)
        tr = np.empty(cnt, dtype=np.int32)
        sl = np.empty(cnt, dtype=np.int64)
        _C.host_perm_positions(pool, list(key.as_words()), n, T, off, cnt,
                               tr.ctypes.data, sl.ctypes.data)
        pos = perm.permute(np.arange(off, off + cnt, dtype=np.uint64), key)
        t_ref, s_ref = plan.position_to_trainer(pos)
        assert np.array_equal(tr, t_ref) and np.array_equal(sl, s_ref)


@pytest.mark.parametrize("trainers", [1, 3])
def test_native_cpu_engine_equals_numpy_engine(small_dataset, trainers):
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    files, n = small_dataset
    plan_args = dict(num_trainers=trainers, num_reducers=4, batch_size=500, drop_last=False)
    gold = CpuShuffleEngine(files, plan_args, L.dataframe_layout, 11, native=False)
    fast = CpuShuffleEngine(files, plan_args, L.dataframe_layout, 11, native=True, num_threads=4)
    assert gold.C is None and fast.C is not None
    try:
        for epoch in range(3):
            gb, fb = gold.start_epoch(epoch), fast.start_epoch(epoch)
            for t in range(trainers):
                gb[t].wait(60); fb[t].wait(60)
                assert np.array_equal(gb[t].data, fb[t].data), (epoch, t)
    finally:
        gold.close(); fast.close()


def test_host_scatter_skips_trainers_not_served():
    """dst == 0 for a trainer means "not mine": rows for it are left alone."""
    pool = _C.HostPool(2)
    n, T, pitch = 5000, 2, 16
    key = perm.make_key(n, 5, 1)
    packed = np.arange(n * pitch, dtype=np.uint8).reshape(n, pitch)
    plan = ShufflePlan(n, T, T, 10)
    out1 = np.zeros((plan.trainer_rows(1), pitch), dtype=np.uint8)
    _C.host_scatter_rows(pool, list(key.as_words()), n, T, packed.ctypes.data, pitch, 0, n,
                         [0, out1.ctypes.data])
    tr, sl = plan.position_to_trainer(perm.permute(np.arange(n, dtype=np.uint64), key))
    want = np.zeros_like(out1)
    want[sl[tr == 1]] = packed[tr == 1]
    assert np.array_equal(out1, want)


def test_native_pack_rows_all_casts_match_numpy_golden():
    """Every (src, dst) cast the packed-row layout supports, list columns and
    odd field alignment: C++ pack == ops/layout.py::pack_rows."""
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import native_pack_rows
    rng = np.random.default_rng(0)
    n = 10_007
    cols = {
        "i64": rng.integers(-(1 << 40), 1 << 40, n, dtype=np.int64),
        "i32": rng.integers(-(1 << 20), 1 << 20, n, dtype=np.int32),
        "u8": rng.integers(0, 255, n, dtype=np.uint8),
        "f32": (rng.random(n, dtype=np.float32) - 0.5) * 1000,
        "f64": (rng.random(n) - 0.5) * 1e6,
        "f16": rng.random(n).astype(np.float16),
        "b": rng.integers(0, 2, n).astype(np.bool_),
        "img": rng.random((n, 5), dtype=np.float32),
    }
    codes = {"i64": L.DT_I64, "i32": L.DT_I32, "u8": L.DT_U8, "f32": L.DT_F32, "f64": L.DT_F64,
             "f16": L.DT_F16, "b": L.DT_BOOL, "img": L.DT_F32}
    spec = []
    for name, src in codes.items():
        for dst in (L.DT_F32, L.DT_BF16, L.DT_F16, L.DT_F64, L.DT_I64, L.DT_I32, L.DT_I16,
                    L.DT_U8, L.DT_BOOL):
            if name in ("f32", "f64", "f16", "img") and dst in (L.DT_I16, L.DT_U8):
                continue                      # out-of-range float -> small int is UB in both
            spec.append((name, src, dst, 5 if name == "img" else 1))
    layout = L.build_layout(spec)
    pool = _C.HostPool(4)
    got = native_pack_rows(_C, pool, cols, layout)
    want = L.pack_rows(cols, layout)
    for f in layout.fields:
        a, b = L.unpack_field(got, f), L.unpack_field(want, f)
        assert np.array_equal(a, b, equal_nan=False) or np.array_equal(
            a.view(np.uint8), b.view(np.uint8)), (f.name, f.src_code, f.dst_code)
    assert np.array_equal(got, want)          # padding bytes are zero in both


def test_perm_cuh_matches_numpy_property():
    """Property test of the shared bijection (csrc/perm.cuh, compiled for the
    host here and for sm_100a in the kernels) against ops/perm.py over random
    table sizes - including > 2^32 rows, where both sides take their 64-bit
    paths - seeds, epochs, trainer counts and source offsets."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    pool = _C.HostPool(2)

    @settings(max_examples=60, deadline=None)
    @given(n=st.one_of(st.integers(1, 5000), st.integers(1 << 20, 1 << 22),
                       st.integers((1 << 32) - 5, (1 << 32) + 5), st.integers(1 << 40, (1 << 40) + 99)),
           seed=st.integers(0, 2**63 - 1), epoch=st.integers(0, 1000),
           T=st.integers(1, 64), frac=st.floats(0, 1))
    def check(n, seed, epoch, T, frac):
        T = min(T, n)
        cnt = min(n, 257)
        off = int((n - cnt) * frac)
        key = perm.make_key(n, seed, epoch)
        plan = ShufflePlan(n, T, T, 10)
        tr = np.empty(cnt, dtype=np.int32)
        sl = np.empty(cnt, dtype=np.int64)
        _C.host_perm_positions(pool, list(key.as_words()), n, T, off, cnt,
                               tr.ctypes.data, sl.ctypes.data)
        pos = perm.permute(np.arange(off, off + cnt, dtype=np.uint64), key)
        assert pos.max() < n
        t_ref, s_ref = plan.position_to_trainer(pos)
        assert np.array_equal(tr, t_ref) and np.array_equal(sl, s_ref)
        # and it really is a bijection on the sampled range
        assert np.array_equal(perm.inverse(pos, key), np.arange(off, off + cnt, dtype=np.uint64))

    check()


def test_cpu_engine_recycles_released_buffers(small_dataset):
    """pandas output: a released epoch buffer is reused by a later epoch (no
    fresh multi-GB allocation per epoch) and the data stays exactly-once."""
    from ray_shuffling_data_loader_b200 import ShufflingDataset
    files, n = small_dataset
    ds = ShufflingDataset(files, 5, 1, 1000, 0, num_reducers=2, backend="cpu", seed=4)
    eng = ds._engine
    assert eng._recycle
    seen_ptrs = []
    for epoch in range(5):
        ds.set_epoch(epoch)
        keys = []
        for df in ds:
            keys.append(df["key"].to_numpy())
        assert np.array_equal(np.sort(np.concatenate(keys)), np.arange(n))
        seen_ptrs.append(sum(len(v) for v in eng._free.values()))
    assert max(seen_ptrs) >= 1                       # something came back to the pool
    # torch output hands out views: never recycled
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    tds = TorchShufflingDataset(files, 1, 1, 1000, 0, num_reducers=2, backend="cpu",
                                feature_columns=["embeddings_name0"], label_column="labels")
    assert not tds.dataset._engine._recycle
    tds.set_epoch(0)
    for _ in tds:
        pass


@pytest.mark.parametrize("native", [True, False])
def test_cpu_engine_delivers_chunks_in_order_and_matches_forward_scatter(small_dataset, native):
    """K7 on the host backend: the epoch is produced reducer chunk by reducer chunk through
    the inverse permutation. (a) waiting for a chunk's rows only returns a complete prefix,
    (b) the result equals an independent forward scatter pi_e(i) -> (trainer, slot)."""
    from ray_shuffling_data_loader_b200.ops import layout as L
    from ray_shuffling_data_loader_b200.ops import perm
    from ray_shuffling_data_loader_b200.runtime import ingest
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    files, n = small_dataset
    plan_args = dict(num_trainers=3, num_reducers=11, batch_size=500, drop_last=False)
    eng = CpuShuffleEngine(files, plan_args, L.dataframe_layout, 21, native=native)
    try:
        # independent golden: decode, pack, forward scatter with the public permutation
        idx = ingest.scan_files(files)
        lay = L.dataframe_layout(idx.schema)
        table = ingest.load_table(idx, 0, n, columns=lay.names)
        packed = L.pack_rows(table.columns, lay)
        for epoch in (0, 1):
            key = perm.make_key(n, 21, epoch)
            pos = perm.permute(np.arange(n, dtype=np.uint64), key)
            trainer, slot = eng.plan.position_to_trainer(pos)
            bufs = eng.start_epoch(epoch)
            seen = {t: [] for t in bufs}
            for t, buf in bufs.items():
                want = np.zeros_like(buf.data)
                sel = trainer == t
                want[slot[sel]] = packed[sel]
                for a, b in eng.plan.trainer_chunks(t):
                    buf.wait(60, row_stop=b)                  # this chunk (and all before it)
                    assert buf._rows_ready >= b
                    seen[t].append(buf._rows_ready)
                    assert np.array_equal(buf.data[:b], want[:b]), (epoch, t, a, b)
                buf.wait(60)
                assert np.array_equal(buf.data, want)
                buf.release()
            assert all(v == sorted(v) for v in seen.values())
    finally:
        eng.close()
