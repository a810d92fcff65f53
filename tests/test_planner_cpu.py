"""The device engine's kernel planner (which scatter kernel writes which bytes
of a packed row) is pure Python: exercise it without a GPU by building the
engine object without running its constructor."""
import numpy as np
import pytest

from ray_shuffling_data_loader_b200.ops import layout as L
from ray_shuffling_data_loader_b200.runtime import device_engine as DE


def _plan(layout, force_generic=False):
    eng = object.__new__(DE.DeviceShuffleEngine)
    eng.layout = layout
    eng.src_fields = list(layout.fields)
    eng.force_generic = force_generic
    eng._plan_kernels()
    return eng


def _cover(eng):
    """Byte ranges claimed by the three kernels; they must not overlap."""
    ranges = []
    if eng.fast_mode >= 0:
        ranges.append((0, eng.fast_write_end))
        if eng.tail_range[1] > eng.tail_range[0]:
            assert eng.tail_range[0] >= eng.fast_write_end      # tail step starts behind the prefix
            ranges.append(eng.tail_range)
    for i in eng.wide_field_idx:
        f = eng.src_fields[i]
        ranges.append((f.offset, f.offset + f.dst_bytes))
    for _, lo, hi in eng.generic_runs:
        ranges.append((lo, hi))
    ranges.sort()
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b <= c, ranges
    return ranges


def test_f32_table_is_one_tma_launch_and_owns_the_padding():
    lay = L.build_layout([(f"f{i}", L.DT_F32, L.DT_F32, 1) for i in range(17)])
    eng = _plan(lay)
    assert eng.fast_mode == 0 and len(eng.fast_field_idx) == 17
    assert not eng.generic_runs and not eng.wide_field_idx
    assert lay.row_pitch == 96 and eng.fast_write_end == 96      # zero-fills bytes 68..95


@pytest.mark.parametrize("dst,mode", [(L.DT_BF16, 1), (L.DT_FP8, 2)])
def test_f32_casts(dst, mode):
    lay = L.build_layout([(f"f{i}", L.DT_F32, dst, 1) for i in range(64)],
                         fp8_block_scale=(dst == L.DT_FP8))
    eng = _plan(lay)
    assert eng.fast_mode == mode
    if mode == 2:
        assert eng.fast_write_end == 64          # payload only: the scales follow
    _cover(eng)


def test_data_spec_native_and_torch_defaults():
    from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
    schema = {"key": (L.DT_I64, 1)}
    schema.update({c: (L.code_from_numpy(dt), 1) for c, (_, _, dt) in DATA_SPEC.items()})
    native = _plan(L.dataframe_layout(schema))
    assert native.fast_mode == 3 and len(native.fast_field_idx) == 21
    assert native.fast_write_end == native.layout.row_pitch == 192
    feats = [c for c in schema if c != "key"]
    torchy = _plan(L.build_layout([(c, schema[c][0], L.DT_F32, 1) for c in feats]))
    assert torchy.fast_mode == 4
    assert torchy.fast_kinds == [0] * 19 + [1]                  # int64 -> f32 ..., float64 -> f32
    assert torchy.fast_write_end == torchy.layout.row_pitch == 96


def test_mixed_prefix_tail_and_kinds():
    cols = [("a", L.DT_I64, L.DT_F32, 1), ("b", L.DT_F64, L.DT_F32, 1),
            ("c", L.DT_I64, L.DT_I32, 1), ("d", L.DT_I64, L.DT_F32, 1),
            ("e", L.DT_I64, L.DT_F32, 1),
            ("key", L.DT_I64, L.DT_I64, 1), ("flag", L.DT_BOOL, L.DT_U8, 1)]
    eng = _plan(L.build_layout(cols))
    # 5 x 4 B then an int64 at byte 24: the TMA kernel keeps the 4 fields of the
    # full 16-byte group, field "e" joins the generic tail
    assert eng.fast_mode == 4 and eng.fast_kinds == [0, 1, 2, 0]
    assert eng.fast_write_end == 16
    assert [eng.src_fields[i].name for i in eng.generic_field_idx] == ["e", "key", "flag"]
    _cover(eng)
    # a tail that starts on the group boundary leaves the prefix whole
    cols8 = cols[:5] + [("f", L.DT_I64, L.DT_F32, 1), ("g", L.DT_I64, L.DT_F32, 1),
                        ("h", L.DT_I64, L.DT_F32, 1)] + cols[5:]
    eng8 = _plan(L.build_layout(cols8))
    assert eng8.fast_mode == 4 and len(eng8.fast_field_idx) == 8 and eng8.fast_write_end == 32
    _cover(eng8)


def test_short_or_unsupported_prefixes_fall_back_to_generic():
    three = _plan(L.build_layout([(f"f{i}", L.DT_F32, L.DT_F32, 1) for i in range(3)]))
    assert three.fast_mode == -1 and len(three.generic_runs) == 1
    assert three.generic_runs[0][1:] == (0, three.layout.row_pitch)
    i16 = _plan(L.build_layout([(f"h{i}", L.DT_I16, L.DT_F32, 1) for i in range(8)]))
    assert i16.fast_mode == -1
    to_bf16 = _plan(L.build_layout([(f"x{i}", L.DT_I64, L.DT_BF16, 1) for i in range(8)]))
    assert to_bf16.fast_mode == -1                              # no 8 -> 2 byte TMA mode
    forced = _plan(L.build_layout([(f"f{i}", L.DT_F32, L.DT_F32, 1) for i in range(64)]),
                   force_generic=True)
    assert forced.fast_mode == -1


def test_image_column_takes_the_wide_kernel():
    lay = L.build_layout([("image", L.DT_F32, L.DT_BF16, 3 * 8 * 8), ("labels", L.DT_I64, L.DT_I64, 1),
                          ("w", L.DT_F32, L.DT_F32, 1)])
    eng = _plan(lay)
    assert eng.wide_field_idx == [0] and eng.fast_mode == -1
    assert [eng.src_fields[i].name for i in eng.generic_field_idx] == ["labels", "w"]
    _cover(eng)


def test_row_align_extends_the_zero_filled_tail():
    lay = L.build_layout([(f"c{i}", L.DT_I64, L.DT_F32, 1) for i in range(21)], row_align=128)
    eng = _plan(lay)
    assert eng.fast_mode == 4 and eng.fast_write_end == 128


def test_planner_invariants_property():
    """Random layouts: every field is written by exactly one kernel, no two
    kernels claim the same byte, TMA fields sit inside the TMA byte range."""
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    casts = {L.DT_F32: [L.DT_F32, L.DT_BF16, L.DT_F16],
             L.DT_I64: [L.DT_I64, L.DT_F32, L.DT_I32, L.DT_BF16],
             L.DT_F64: [L.DT_F64, L.DT_F32], L.DT_I32: [L.DT_I32, L.DT_F32],
             L.DT_U8: [L.DT_U8, L.DT_F32], L.DT_I16: [L.DT_I16]}
    field = st.sampled_from(list(casts)).flatmap(
        lambda src: st.tuples(st.just(src), st.sampled_from(casts[src]),
                              st.sampled_from([1, 1, 1, 1, 1, 3, 16, 48])))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(field, min_size=1, max_size=40), st.sampled_from([0, 0, 64, 128]))
    def check(spec, align):
        cols = [(f"c{i}", s, d, w) for i, (s, d, w) in enumerate(spec)]
        lay = L.build_layout(cols, row_align=align)
        try:
            eng = _plan(lay)
        except ValueError as e:           # the documented refusal, never a wrong plan
            assert "overlap" in str(e)
            return
        claimed = sorted(eng.fast_field_idx + eng.tail_field_idx + eng.wide_field_idx
                         + eng.generic_field_idx)
        assert claimed == list(range(len(lay.fields)))
        for i in eng.tail_field_idx:                 # tails: small scalars inside the tail range
            f = lay.fields[i]
            assert f.width == 1 and L.itemsize(f.dst_code) in (4, 8)
            assert eng.tail_range[0] <= f.offset and f.offset + f.dst_bytes <= eng.tail_range[1]
        assert len(eng.tail_field_idx) <= 4 and not (eng.tail_field_idx and eng.generic_field_idx)
        ranges = _cover(eng)
        assert all(b <= lay.row_pitch for _, b in ranges)
        for i in eng.fast_field_idx:
            f = lay.fields[i]
            assert f.offset + f.dst_bytes <= eng.fast_write_end
        for idxs, lo, hi in eng.generic_runs:
            for i in idxs:
                f = lay.fields[i]
                assert lo <= f.offset and f.offset + f.dst_bytes <= hi
        if eng.fast_mode in (3, 4):
            assert all(L.itemsize(lay.fields[i].src_code) == 8 for i in eng.fast_field_idx)
        if eng.fast_mode == 4:
            assert len(eng.fast_kinds) == len(eng.fast_field_idx)

    check()


def test_dataframe_layout_stores_the_tma_class_first():
    """An int64 key in front of 64 float32 columns: stored behind them so the TMA
    kernel takes the floats; user-facing column order stays the file's."""
    schema = {"key": (L.DT_I64, 1)}
    schema.update({f"f{i}": (L.DT_F32, 1) for i in range(63)})
    schema["labels"] = (L.DT_F32, 1)
    lay = L.dataframe_layout(schema)
    assert [f.name for f in lay.fields[:3]] == ["f0", "f1", "f2"] and lay.fields[-1].name == "key"
    assert lay.names == list(schema)                              # display order
    assert lay.field("key").offset == 256 and lay.row_pitch == 288
    eng = _plan(lay)
    assert eng.fast_mode == 0 and len(eng.fast_field_idx) == 64
    # the key rides the fast kernel as a tail field: no second (generic) launch
    assert [eng.src_fields[i].name for i in eng.tail_field_idx] == ["key"]
    assert not eng.generic_runs and eng.tail_range == (256, 288)
    _cover(eng)
    # natural order already optimal (DATA_SPEC + key: all 8-byte) -> untouched
    from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
    s2 = {"key": (L.DT_I64, 1)}
    s2.update({c: (L.code_from_numpy(dt), 1) for c, (_, _, dt) in DATA_SPEC.items()})
    lay2 = L.dataframe_layout(s2)
    assert lay2.display_order is None and [f.name for f in lay2.fields] == list(s2)
    # opt-out
    assert [f.name for f in L.dataframe_layout(schema, optimize=False).fields] == list(schema)
    # round trip through pack / DataFrame keeps names and values
    from ray_shuffling_data_loader_b200.runtime.chunks import packed_to_dataframe
    data = {n: (np.arange(50, dtype=np.int64) if n == "key"
                else np.full(50, i, dtype=np.float32)) for i, n in enumerate(schema)}
    df = packed_to_dataframe(L.pack_rows(data, lay), lay)
    assert list(df.columns) == list(schema)
    assert np.array_equal(df["key"].to_numpy(), np.arange(50)) and df["f5"].iloc[0] == 6.0


def test_torch_layout_reorders_storage_not_tensors(small_dataset):
    """An int64 id (kept as int64) in front of float features: stored behind them;
    the yielded feature list keeps the caller's order and values."""
    torch = pytest.importorskip("torch")
    from ray_shuffling_data_loader_b200 import TorchShufflingDataset
    files, n = small_dataset
    feats = ["key"] + [f"embeddings_name{i}" for i in range(6)]
    types = [torch.int64] + [torch.float32] * 6
    ds = TorchShufflingDataset(files, 1, 1, 1000, 0, num_reducers=2, backend="cpu", seed=2,
                               feature_columns=feats, feature_types=types, label_column="labels")
    lay = ds.dataset._engine.layout
    assert lay.fields[0].name == "embeddings_name0" and lay.fields[-1].name == "key"
    eng = _plan(lay)
    assert eng.fast_mode == 4 and len(eng.fast_field_idx) == 7       # 6 features + label
    ds.set_epoch(0)
    keys = []
    for features, label in ds:
        assert len(features) == 7 and features[0].dtype == torch.int64
        assert features[1].dtype == torch.float32 and features[1].shape[1] == 1
        keys.append(features[0][:, 0].clone())
    assert sorted(torch.cat(keys).tolist()) == list(range(n))


def test_default_chunk_passes_and_pass_of_row():
    """K7 planning is pure Python: how many destination-chunk passes are free, and
    which pass completes a given row range."""
    from ray_shuffling_data_loader_b200.ops.plan import balanced_split
    f32 = L.build_layout([(f"f{i}", L.DT_F32, L.DT_F32, 1) for i in range(64)])
    assert DE.default_chunk_passes(f32, 1, 8) == 1            # one GPU: a pass is never free
    assert DE.default_chunk_passes(f32, 2, 8) == 1            # 256 B rows, half of them remote
    assert DE.default_chunk_passes(f32, 4, 8) == 2
    assert DE.default_chunk_passes(f32, 8, 8) == 2
    assert DE.default_chunk_passes(f32, 8, 1) == 1            # never more than the reducer chunks
    bf16 = L.build_layout([(f"f{i}", L.DT_F32, L.DT_BF16, 1) for i in range(64)])
    assert DE.default_chunk_passes(bf16, 8, 8) == 1           # half the bytes on the wire
    eng = object.__new__(DE.DeviceShuffleEngine)
    eng.chunk_passes = 3
    eng.pass_bounds = balanced_split(10, 3)                   # [0,4) [4,7) [7,10)
    assert [eng.pass_of_row(r) for r in (1, 4, 5, 7, 8, 10)] == [0, 0, 1, 1, 2, 2]
    assert eng.pass_of_row(0) == 0
