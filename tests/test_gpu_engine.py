"""GPU tests (run under gpurun): every CUDA kernel is diffed byte for byte
against the numpy golden engine (implication 3)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.ops import perm
from ray_shuffling_data_loader_b200.ops.plan import ShufflePlan


def _native():
    from ray_shuffling_data_loader_b200.runtime.device_engine import load_native
    return load_native()


def _engines(filenames, layout_fn, num_trainers, seed=5, batch_size=512, **dev_opts):
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    plan_args = dict(num_trainers=num_trainers, num_reducers=max(2, num_trainers),
                     batch_size=batch_size, drop_last=False)
    cpu = CpuShuffleEngine(filenames, plan_args, layout_fn, seed)
    dev = DeviceShuffleEngine(filenames, plan_args, layout_fn, seed, **dev_opts)
    return cpu, dev


def _compare_epochs(cpu, dev, epochs=(0, 1, 2), fields_only=False):
    for epoch in epochs:
        cb = cpu.start_epoch(epoch)
        db = dev.start_epoch(epoch)
        for t in dev.local_trainers:
            cb[t].wait(60)
            db[t].wait(60)
            got = db[t].data.cpu().numpy()
            want = cb[t].data
            assert got.shape == want.shape
            if fields_only:
                for f in dev.layout.fields:
                    assert np.array_equal(L.unpack_field(got, f), L.unpack_field(want, f)), \
                        (epoch, t, f.name)
            else:
                assert np.array_equal(got, want), (epoch, t)
        for t in dev.local_trainers:
            db[t].release()
    torch.cuda.synchronize()


def _float_files(tmp_path_factory, ncols, nrows=20_011, nfiles=3, name="f"):
    from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec
    d = tmp_path_factory.mktemp(f"{name}{ncols}")
    files, _ = generate_data(nrows, nfiles, 2, 0.0, str(d),
                             data_spec=float_spec(ncols, np.float32), seed=ncols)
    return files


def _f32_layout(cols, dst=L.DT_F32, fp8=False):
    def fn(schema):
        return L.build_layout([(c, schema[c][0], dst, 1) for c in cols], fp8_block_scale=fp8)
    return fn


def test_perm_positions_matches_numpy():
    C = _native()
    for n, T, off, cnt in [(1000, 1, 0, 1000), (100_003, 8, 12_345, 40_000),
                           (7, 3, 0, 7), (1 << 20, 5, 999, 5000)]:
        for epoch in (0, 3):
            key = perm.make_key(n, 77, epoch)
            plan = ShufflePlan(n, T, T, 10)
            tr = torch.empty(cnt, dtype=torch.int32, device="cuda")
            sl = torch.empty(cnt, dtype=torch.int64, device="cuda")
            C.perm_positions(list(key.as_words()), n, T, off, cnt, tr.data_ptr(),
                             sl.data_ptr(), torch.cuda.current_stream().cuda_stream)
            pos = perm.permute(np.arange(off, off + cnt, dtype=np.uint64), key)
            t_ref, s_ref = plan.position_to_trainer(pos)
            assert np.array_equal(tr.cpu().numpy(), t_ref)
            assert np.array_equal(sl.cpu().numpy(), s_ref)


@pytest.mark.parametrize("ncols,trainers", [(64, 1), (64, 3), (17, 2), (63, 1), (200, 2), (4, 1)])
def test_fast_f32_scatter_matches_golden(tmp_path_factory, ncols, trainers):
    files = _float_files(tmp_path_factory, ncols)
    cols = [f"f{i}" for i in range(ncols - 1)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), trainers)
    assert dev.fast_mode == 0 and not dev.generic_field_idx
    try:
        _compare_epochs(cpu, dev)
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("ncols", [64, 40, 130])
def test_fast_bf16_scatter_matches_golden(tmp_path_factory, ncols):
    files = _float_files(tmp_path_factory, ncols, name="b")
    cols = [f"f{i}" for i in range(ncols - 1)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols, L.DT_BF16), 2)
    assert dev.fast_mode == 1
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("ncols", [64, 128, 100, 300])
def test_fast_fp8_block_scaled_matches_golden(tmp_path_factory, ncols):
    files = _float_files(tmp_path_factory, ncols, name="q")
    cols = [f"f{i}" for i in range(ncols - 1)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols, L.DT_FP8, fp8=True), 2)
    assert dev.fast_mode == 2
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("tails", [True, False])
def test_split_fast_features_generic_label(tmp_path_factory, tails):
    """bf16 features + a float32 label + an int64 key: the two trailing scalars ride the
    fast kernel as tail fields (one launch), or - tail_fields=False - the generic kernel."""
    files = _float_files(tmp_path_factory, 65, name=f"s{int(tails)}")
    feats = [f"f{i}" for i in range(64)]

    def fn(schema):
        return L.build_layout([(c, L.DT_F32, L.DT_BF16, 1) for c in feats]
                              + [("labels", L.DT_F32, L.DT_F32, 1), ("key", L.DT_I64, L.DT_I64, 1)])
    cpu, dev = _engines(files, fn, 2, tail_fields=tails)
    assert dev.fast_mode == 1
    assert (len(dev.tail_field_idx), len(dev.generic_field_idx)) == ((2, 0) if tails else (0, 2))
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


def test_generic_data_spec_matches_golden(small_dataset):
    files, n = small_dataset
    cpu, dev = _engines(files, L.dataframe_layout, 3, force_generic=True)
    assert dev.fast_mode == -1
    try:
        _compare_epochs(cpu, dev)
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("trainers,resident", [(1, "hbm"), (3, "hbm"), (2, "host")])
def test_typed64_copy_data_spec_matches_golden(small_dataset, trainers, resident):
    """DATA_SPEC rows in their native dtypes (21 x int64/float64 = what plain
    ShufflingDataset shuffles): one TMA launch, mode 3 (8-byte bit copy)."""
    files, n = small_dataset
    opts = dict(resident=resident)
    if resident == "host":
        opts["stream_chunk_rows"] = 2500
    cpu, dev = _engines(files, L.dataframe_layout, trainers, **opts)
    assert dev.fast_mode == 3 and not dev.generic_field_idx
    try:
        _compare_epochs(cpu, dev)
    finally:
        dev.close(); cpu.close()


def _int_files(tmp_path_factory, ncols, nrows=9_973):
    from ray_shuffling_data_loader_b200.data_generation import generate_data
    d = tmp_path_factory.mktemp(f"i{ncols}")
    spec = {f"c{i}": ((-(1 << 40), 1 << 40, np.int64) if i % 3 else (0, 1, np.float64))
            for i in range(ncols)}
    files, _ = generate_data(nrows, 2, 2, 0.0, str(d), data_spec=spec, seed=ncols)
    return files, list(spec)


@pytest.mark.parametrize("ncols,trainers", [(21, 2), (70, 1), (33, 3), (5, 2)])
def test_typed64_convert_matches_golden(tmp_path_factory, ncols, trainers):
    """8-byte sources cast to 4-byte fields in the TMA kernel (mode 4):
    int64 -> f32 (values beyond 2^24: round to nearest even), float64 -> f32,
    int64 -> int32; more than 32 columns exercises the panel loop."""
    files, names = _int_files(tmp_path_factory, ncols)

    def fn(schema):
        cols = []
        for i, c in enumerate(names):
            src = schema[c][0]
            dst = L.DT_I32 if (src == L.DT_I64 and i % 5 == 1) else L.DT_F32
            cols.append((c, src, dst, 1))
        return L.build_layout(cols)
    cpu, dev = _engines(files, fn, trainers)
    assert dev.fast_mode == 4 and not dev.generic_field_idx
    assert len(dev.fast_kinds) == ncols and set(dev.fast_kinds) <= {0, 1, 2}
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("tails,passes", [(True, 1), (True, 3), (False, 1)])
def test_typed64_prefix_then_generic_tail(small_dataset, tails, passes):
    """The torch default on DATA_SPEC plus an int64 key kept as int64: the float32
    prefix rides mode 4, the 8-byte key is a tail field (or the generic kernel's)."""
    files, n = small_dataset

    def fn(schema):
        feats = [c for c in schema if c not in ("key",)]
        return L.build_layout([(c, schema[c][0], L.DT_F32, 1) for c in feats]
                              + [("key", L.DT_I64, L.DT_I64, 1)])
    cpu, dev = _engines(files, fn, 2, tail_fields=tails, chunk_passes=passes)
    assert dev.fast_mode == 4
    assert (len(dev.tail_field_idx), len(dev.generic_field_idx)) == ((1, 0) if tails else (0, 1))
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("tails,resident", [(True, "hbm"), (True, "host"), (False, "hbm")])
def test_prefix_trimmed_to_whole_16_byte_groups(small_dataset, tails, resident):
    """5 x (int64 -> f32) followed by an int64 at byte 24: the TMA kernel takes
    the 4 fields of the full 16-byte group, the generic kernel the rest."""
    files, n = small_dataset

    def fn(schema):
        return L.build_layout([(f"embeddings_name{i}", L.DT_I64, L.DT_F32, 1) for i in range(5)]
                              + [("key", L.DT_I64, L.DT_I64, 1), ("labels", L.DT_F64, L.DT_F32, 1)])
    opts = dict(tail_fields=tails, resident=resident)
    if resident == "host":
        opts["stream_chunk_rows"] = 1024
    cpu, dev = _engines(files, fn, 2, **opts)
    assert dev.fast_mode == 4 and len(dev.fast_field_idx) == 4 and dev.fast_write_end == 16
    # embeddings_name4 (int64 -> f32), key (int64) and labels (float64 -> f32) follow
    assert (len(dev.tail_field_idx), len(dev.generic_field_idx)) == ((3, 0) if tails else (0, 3))
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


def test_generic_casts_match_golden(small_dataset):
    files, n = small_dataset

    def fn(schema):
        return L.build_layout([
            ("embeddings_name0", L.DT_I64, L.DT_F32, 1), ("embeddings_name12", L.DT_I64, L.DT_I32, 1),
            ("labels", L.DT_F64, L.DT_F32, 1), ("labels", L.DT_F64, L.DT_BF16, 1),
            ("one_hot0", L.DT_I64, L.DT_U8, 1), ("labels", L.DT_F64, L.DT_F16, 1),
            ("embeddings_name3", L.DT_I64, L.DT_F64, 1), ("key", L.DT_I64, L.DT_I64, 1),
            ("one_hot1", L.DT_I64, L.DT_BOOL, 1), ("embeddings_name1", L.DT_I64, L.DT_I16, 1)])
    cpu, dev = _engines(files, fn, 2)
    try:
        _compare_epochs(cpu, dev, epochs=(0,))
    finally:
        dev.close(); cpu.close()


def test_one_d_bulk_path_matches_golden(tmp_path_factory):
    """The three source-load paths - 1-D bulk copies (UBLKCP, default), swizzled
    tensor-map boxes and one dense tensor-map box (UTMALDG) - must all agree
    with the golden."""
    files = _float_files(tmp_path_factory, 70, name="o")
    cols = [f"f{i}" for i in range(69)] + ["labels"]
    for mode in (0, 1, 2):
        cpu, dev = _engines(files, _f32_layout(cols), 2)
        dev.tmap_mode = mode
        assert dev.fast_mode == 0
        try:
            _compare_epochs(cpu, dev, epochs=(0, 1))
        finally:
            dev.close(); cpu.close()


def test_forced_generic_equals_fast(tmp_path_factory):
    files = _float_files(tmp_path_factory, 64, name="g")
    cols = [f"f{i}" for i in range(63)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 2, force_generic=True)
    assert dev.fast_mode == -1
    try:
        _compare_epochs(cpu, dev, epochs=(0,))
    finally:
        dev.close(); cpu.close()


def test_streaming_host_resident_matches_golden(tmp_path_factory):
    files = _float_files(tmp_path_factory, 32, name="h")
    cols = [f"f{i}" for i in range(31)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 2, resident="host", stream_chunk_rows=3000)
    assert dev.h2d_bytes_per_epoch() == 20_011 * 32 * 4
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1, 2, 3))
    finally:
        dev.close(); cpu.close()


def test_nccl_baseline_single_gpu_matches_golden(tmp_path_factory):
    files = _float_files(tmp_path_factory, 64, name="n")
    cols = [f"f{i}" for i in range(63)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 2, exchange="nccl")
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


def test_stream_wait_mode(tmp_path_factory):
    files = _float_files(tmp_path_factory, 64, name="w")
    cols = [f"f{i}" for i in range(63)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 1, wait_mode="stream")
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1, 2))
        dev.check_error()
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("sched", [0, 1])
def test_producer_schedules_agree(tmp_path_factory, small_dataset, sched):
    """Both producer schedules of the TMA kernel (loader warps + tile-per-warp
    index vs cooperative index warps that also load) must give the golden bytes,
    for 4-byte and 8-byte sources, single- and multi-panel tables."""
    files = _float_files(tmp_path_factory, 150, name=f"sc{sched}")
    cols = [f"f{i}" for i in range(149)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 3, sched=sched)
    assert dev.sched == sched
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()
    dfiles, _ = small_dataset
    cpu, dev = _engines(dfiles, L.dataframe_layout, 2, sched=sched)
    assert dev.fast_mode == 3
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("mode", ["stream", "host"])
def test_backpressure_modes(tmp_path_factory, mode):
    """Slot reuse beyond the window: the consumed-flag gate runs as a wait
    kernel on the shuffle stream (default) or as a host poll."""
    files = _float_files(tmp_path_factory, 64, name="bp" + mode)
    cols = [f"f{i}" for i in range(63)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 2, backpressure=mode)
    assert dev.backpressure == mode
    try:
        _compare_epochs(cpu, dev, epochs=tuple(range(6)))
        dev.check_error()
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("mode", ["stream", "host"])
def test_stalled_consumer_is_reported_not_hung(tmp_path_factory, mode):
    """Fault injection (the reference deadlocks on a dead trainer):
    a trainer that never releases its epoch must surface as a TimeoutError
    within flag_timeout_s - from the host poll, or from the error word a
    device-side wait kernel leaves behind."""
    files = _float_files(tmp_path_factory, 16, name="fi" + mode)
    cols = [f"f{i}" for i in range(15)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 1, backpressure=mode, flag_timeout_s=0.5)
    try:
        b0 = dev.start_epoch(0)
        b1 = dev.start_epoch(1)
        b0[0].wait(30); b1[0].wait(30)          # both slots full, nothing released
        if mode == "host":
            with pytest.raises(TimeoutError, match="did not release epoch 0"):
                dev.start_epoch(2)
        else:
            dev.start_epoch(2)                   # enqueues the wait kernel; gives up after 0.5 s
            torch.cuda.synchronize()
            with pytest.raises(TimeoutError, match="device-side flag wait timed out"):
                dev.check_error()
    finally:
        dev.close(); cpu.close()


def _image_files(tmp_path_factory, n=3001, px=3 * 16 * 16):
    import pyarrow as pa
    import pyarrow.parquet as pq
    d = tmp_path_factory.mktemp("img")
    rng = np.random.default_rng(0)
    files = []
    for i, (a, b) in enumerate([(0, n // 3), (n // 3, n)]):
        img = rng.random((b - a, px), dtype=np.float32)
        tbl = pa.table({"image": pa.FixedSizeListArray.from_arrays(pa.array(img.reshape(-1)), px),
                        "labels": pa.array(np.arange(a, b, dtype=np.int64)),
                        "w": pa.array(rng.random(b - a).astype(np.float32))})
        fn = str(d / f"img{i}.parquet")
        pq.write_table(tbl, fn, row_group_size=500)
        files.append(fn)
    return files, px


@pytest.mark.parametrize("dst", [L.DT_F32, L.DT_BF16, L.DT_F16])
def test_wide_list_column_matches_golden(tmp_path_factory, dst):
    """List-valued (image) columns take the row-major wide kernel; the int64
    label and a float scalar ride the generic kernel next to it."""
    files, px = _image_files(tmp_path_factory)

    def fn(schema):
        return L.build_layout([("image", L.DT_F32, dst, px), ("labels", L.DT_I64, L.DT_I64, 1),
                               ("w", L.DT_F32, L.DT_F32, 1)])
    cpu, dev = _engines(files, fn, 2)
    assert dev.wide_field_idx == [0] and dev.fast_mode == -1 and len(dev.generic_runs) == 1
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


def test_wide_column_streaming_and_torch_api(tmp_path_factory):
    files, px = _image_files(tmp_path_factory)
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    ds = TorchShufflingDataset(files, 2, 1, 256, 0, num_reducers=2, feature_columns=["image"],
                               feature_shapes=[(3, 16, 16)], feature_types=[torch.bfloat16],
                               label_column="labels", label_type=torch.int64, seed=1,
                               queue_name="gw", resident="host", stream_chunk_rows=512)
    for epoch in range(2):
        ds.set_epoch(epoch)
        labels = []
        for (img,), y in ds:
            assert img.is_cuda and img.dtype == torch.bfloat16 and img.shape[1:] == (3, 16, 16)
            labels.append(y[:, 0].clone())
        assert sorted(torch.cat(labels).tolist()) == list(range(3001))


# ---------------------------------------------------------------------------
# K7: destination-chunk passes with their own completion flags
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("passes", [2, 5])
def test_chunk_passes_match_golden_all_kernels(tmp_path_factory, small_dataset, passes):
    """Splitting an epoch's scatter into destination-chunk passes must not change a
    byte: fast f32 (multi-panel), typed-64 copy, generic casts and the wide kernel."""
    files = _float_files(tmp_path_factory, 70, name=f"cp{passes}")
    cols = [f"f{i}" for i in range(69)] + ["labels"]
    cpu, dev = _engines(files, _f32_layout(cols), 3, chunk_passes=passes)
    assert dev.chunk_passes == passes and len(dev.pass_bounds) == passes
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1, 2, 3))
        dev.check_error()
    finally:
        dev.close(); cpu.close()
    dfiles, _ = small_dataset
    cpu, dev = _engines(dfiles, L.dataframe_layout, 2, chunk_passes=passes)
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()
    cpu, dev = _engines(dfiles, L.dataframe_layout, 2, chunk_passes=passes, force_generic=True)
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


def test_chunk0_consumable_while_scatter_still_running(tmp_path_factory):
    """K7 (reference dataset.py:133-139: a trainer starts on the first finished
    reducer output). With chunk passes, chunk 0's flag fires after the first pass:
    a copy of chunk 0 enqueued behind ``chunk.wait()`` on the consumer stream
    (a) holds the golden bytes and (b) finishes before the epoch's last pass does."""
    from ray_shuffling_data_loader_b200.runtime.chunks import ShuffledChunk
    C = _native()
    files = _float_files(tmp_path_factory, 64, nrows=1_500_000, nfiles=4, name="ttfb")
    cols = [f"f{i}" for i in range(63)] + ["labels"]
    plan_reducers = 4
    from ray_shuffling_data_loader_b200.runtime.cpu_engine import CpuShuffleEngine
    from ray_shuffling_data_loader_b200.runtime.device_engine import DeviceShuffleEngine
    plan_args = dict(num_trainers=1, num_reducers=plan_reducers, batch_size=4096, drop_last=False)
    cpu = CpuShuffleEngine(files, plan_args, _f32_layout(cols), 9)
    dev = DeviceShuffleEngine(files, plan_args, _f32_layout(cols), 9, chunk_passes=4,
                              wait_mode="stream")
    try:
        want = cpu.start_epoch(0)[0]
        want.wait(120)
        dev._ensure_ingested(0)
        torch.cuda.synchronize()
        gate = torch.zeros(1, dtype=torch.int32, device="cuda")
        side = torch.cuda.Stream()
        lead = []
        for epoch in (0, 1):
            # Hold the shuffle stream behind a gate flag until the consumer's work is
            # queued too, so the comparison below is about device-side ordering only
            # (not about how fast Python enqueues kernels).
            C.wait_flags(gate.data_ptr(), 1, epoch + 1, int(60e9), dev.arena + dev.off_error,
                         dev.shuffle_stream)
            buf = dev.start_epoch(epoch)[0]
            stream = torch.cuda.current_stream().cuda_stream
            chunks = [ShuffledChunk(buf, i, a, b)
                      for i, (a, b) in enumerate(dev.plan.trainer_chunks(0))]
            assert len(chunks) == 4
            chunks[0].wait()                      # enqueues ONE wait kernel: pass 0's flags
            snap = buf.data[chunks[0].row_start:chunks[0].row_stop].clone()
            e_c0 = C.event_create(True)
            C.event_record(e_c0, stream)
            e_end = C.event_create(True)
            C.event_record(e_end, dev.shuffle_stream)
            C.signal_flags([gate.data_ptr()], epoch + 1, side.cuda_stream)     # open the gate
            torch.cuda.synchronize()
            lead.append(C.event_elapsed_ms(e_c0, e_end))
            if epoch == 0:
                assert np.array_equal(snap.cpu().numpy(),
                                      want.data[chunks[0].row_start:chunks[0].row_stop])
            for c in chunks[1:]:
                c.wait()
            assert dev.first_pass_ms(epoch) is not None
            assert dev.first_pass_ms(epoch) < 0.6 * dev.epoch_kernel_ms(epoch)
            buf.release()
        # chunk 0 (and its copy) was done while later passes were still running
        assert min(lead) > 0.0, lead
        dev.check_error()
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("dst", [L.DT_U8, L.DT_F32, L.DT_BF16, L.DT_F16])
def test_uint8_image_column_matches_golden(tmp_path_factory, dst):
    """uint8 pixels (BASELINE config 5: 3 x 224 x 224 images stored as bytes): the
    wide kernel's 16-byte copy (uint8 out) and its vectorised uint8 -> f32 / bf16 /
    f16 conversion (16 pixels per lane per step) against the numpy golden."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    d = tmp_path_factory.mktemp(f"u8img{dst}")
    px, n = 3 * 32 * 32, 1500
    rng = np.random.default_rng(3)
    files = []
    for i in range(2):
        img = rng.integers(0, 256, (n, px), dtype=np.uint8)
        tbl = pa.table({"image": pa.FixedSizeListArray.from_arrays(pa.array(img.reshape(-1)), px),
                        "labels": pa.array(np.arange(i * n, (i + 1) * n, dtype=np.int64))})
        fn = str(d / f"u8_{i}.parquet")
        pq.write_table(tbl, fn, row_group_size=400)
        files.append(fn)

    def fn_layout(schema):
        assert schema["image"] == (L.DT_U8, px)
        return L.build_layout([("image", L.DT_U8, dst, px), ("labels", L.DT_I64, L.DT_I64, 1)])
    cpu, dev = _engines(files, fn_layout, 2, chunk_passes=2)
    assert dev.wide_field_idx == [0]
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1))
    finally:
        dev.close(); cpu.close()


@pytest.mark.parametrize("schema", ["f32", "dataspec"])
def test_disk_streaming_matches_golden(tmp_path_factory, small_dataset, schema):
    """resident='disk': nothing is kept between epochs - every epoch re-decodes its
    Parquet row groups through a bounded ring of pinned staging slots (tables larger
    than host memory; the reference's only mode, shuffle.py:151)."""
    if schema == "f32":
        files = _float_files(tmp_path_factory, 64, nrows=30_011, nfiles=4, name="disk")
        cols = [f"f{i}" for i in range(63)] + ["labels"]
        layout = _f32_layout(cols)
    else:
        files, _ = small_dataset
        layout = L.dataframe_layout
    cpu, dev = _engines(files, layout, 2, resident="disk", num_threads=3)
    assert dev.chunk_passes == 1 and dev.h2d_bytes_per_epoch() > 0
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1, 2, 3))
        assert dev.h2d_bytes_enqueued == 4 * dev.h2d_bytes_per_epoch()
        assert not getattr(dev, "host_cols", None)          # no resident host table
        dev.check_error()
    finally:
        dev.close(); cpu.close()


def test_destination_offsets_beyond_4_gib():
    """Byte offsets into an epoch buffer exceed 2^32 for the headline table already
    (12.5 M x 256 B = 3.2 GB is close; 8-GPU wide rows are far beyond). A 5.1 GB
    destination, 20 M rows x 64 f32, two trainers: sampled rows - specifically
    including ones that land above the 4 GiB mark - carry the right bytes."""
    C = _native()
    free, _ = torch.cuda.mem_get_info()
    if free < 14 * 2**30:
        pytest.skip("needs ~11 GB of free device memory")
    n, F, pitch = 20_000_000, 64, 256
    rows_pad = (n + 255) // 256 * 256
    src = torch.rand((F, rows_pad), dtype=torch.float32, device="cuda")
    dst = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda")       # one 5.12 GB block
    assert dst.numel() > 2**32
    T = 2
    per = n // T
    ptrs = torch.tensor([src[c].data_ptr() for c in range(F)], dtype=torch.int64, device="cuda")
    key = perm.make_key(n, 77, 3)
    stream = torch.cuda.current_stream().cuda_stream
    C.scatter_fast(key=list(key.as_words()), num_rows=n, num_trainers=T, cols=ptrs.data_ptr(),
                   num_cols=F, n_local=n, global_offset=0, row_pitch=pitch, scale_offset=0,
                   dst=[dst.data_ptr(), dst.data_ptr() + per * pitch], mode=0,
                   grid=C.sm_count(0), stream=stream, write_end=pitch)
    torch.cuda.synchronize()
    # sample source rows from all over the table; their positions are uniformly spread
    idx = np.concatenate([np.arange(0, 50_000), np.arange(n - 50_000, n),
                          np.random.default_rng(0).integers(0, n, 100_000)]).astype(np.uint64)
    pos = perm.permute(idx, key).astype(np.int64)          # trainer t owns [t*per, (t+1)*per)
    assert (pos * pitch > 2**32).sum() > 10_000            # the >4 GiB region is exercised
    got = dst[torch.from_numpy(pos).cuda()].view(torch.float32)
    want = src[:, torch.from_numpy(idx.astype(np.int64)).cuda()].t().contiguous()
    assert torch.equal(got, want)
    # every row written exactly once: the column-0 sums agree
    assert torch.allclose(dst.view(torch.float32)[:, 0].sum(dtype=torch.float64),
                          src[0, :n].sum(dtype=torch.float64), rtol=1e-9)


def test_fp8_features_with_f32_label_and_key_as_tail_fields(tmp_path_factory):
    """Block-scaled fp8 features, a float32 label and an int64 key between the payload and
    the UE8M0 scale bytes: one launch of the fast kernel writes all of it."""
    files = _float_files(tmp_path_factory, 65, name="fp8tail")
    feats = [f"f{i}" for i in range(64)]

    def fn(schema):
        return L.build_layout([(c, L.DT_F32, L.DT_FP8, 1) for c in feats]
                              + [("labels", L.DT_F32, L.DT_F32, 1)], fp8_block_scale=True)
    cpu, dev = _engines(files, fn, 2)
    assert dev.fast_mode == 2 and len(dev.tail_field_idx) == 1 and not dev.generic_runs
    assert dev.tail_range == (64, dev.layout.scale_offset)
    try:
        _compare_epochs(cpu, dev, epochs=(0, 1, 2))
    finally:
        dev.close(); cpu.close()
