"""K8 gather/cast-to-tensor golden: packed layout, bf16 and block-scaled fp8."""
import numpy as np
import pytest
import torch

from ray_shuffling_data_loader_b200.ops import fp8
from ray_shuffling_data_loader_b200.ops import layout as L


def _cols(n=257):
    rng = np.random.default_rng(0)
    return {
        "a": rng.integers(-1000, 1000, n, dtype=np.int64),
        "b": rng.random(n).astype(np.float64),
        "c": rng.random(n).astype(np.float32),
        "d": rng.integers(0, 100, n, dtype=np.int32),
        "e": rng.random((n, 3)).astype(np.float32),
    }


def test_pack_unpack_native():
    cols = _cols()
    schema = {k: (L.code_from_numpy(v.dtype), 1 if v.ndim == 1 else v.shape[1])
              for k, v in cols.items()}
    lay = L.dataframe_layout(schema)
    assert lay.row_pitch % 16 == 0
    for f in lay.fields:
        assert f.offset % L.itemsize(f.dst_code) == 0
    packed = L.pack_rows(cols, lay)
    assert packed.shape == (257, lay.row_pitch)
    for f in lay.fields:
        assert np.array_equal(L.unpack_field(packed, f), cols[f.name])
    idx = np.array([5, 1, 200, 5])
    sub = L.pack_rows(cols, lay, idx)
    assert np.array_equal(sub, packed[idx])


def test_casts_match_torch():
    cols = _cols()
    lay = L.build_layout([("a", L.DT_I64, L.DT_F32, 1), ("b", L.DT_F64, L.DT_F32, 1),
                          ("c", L.DT_F32, L.DT_BF16, 1), ("d", L.DT_I32, L.DT_I64, 1),
                          ("b2", L.DT_F64, L.DT_F16, 1)])
    cols["b2"] = cols["b"]
    packed = L.pack_rows(cols, lay)
    got = {f.name: L.unpack_field(packed, f) for f in lay.fields}
    assert np.array_equal(got["a"], torch.as_tensor(cols["a"]).to(torch.float32).numpy())
    assert np.array_equal(got["b"], torch.as_tensor(cols["b"]).to(torch.float32).numpy())
    bf = torch.as_tensor(cols["c"]).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got["c"], bf)
    assert np.array_equal(got["d"], cols["d"].astype(np.int64))
    assert np.array_equal(got["b2"], cols["b"].astype(np.float16))


def test_fast_path_detection():
    f32 = L.build_layout([(f"f{i}", L.DT_F32, L.DT_F32, 1) for i in range(64)])
    assert f32.uniform_code == L.DT_F32 and f32.is_fast_path and f32.row_pitch == 256
    bf = L.build_layout([(f"f{i}", L.DT_F32, L.DT_BF16, 1) for i in range(64)])
    assert bf.is_fast_path and bf.row_pitch == 128
    mixed = L.build_layout([("a", L.DT_I64, L.DT_F32, 1), ("c", L.DT_F32, L.DT_F32, 1)])
    assert not mixed.is_fast_path


def test_e4m3_roundtrip_all_codes():
    codes = np.arange(256, dtype=np.uint8)
    vals = fp8.e4m3_bits_to_f32(codes)
    finite = ~np.isnan(vals)
    back = fp8.f32_to_e4m3_bits(vals[finite])
    # -0.0 and +0.0 keep their sign bit; every finite code round-trips
    assert np.array_equal(back, codes[finite])
    assert fp8.f32_to_e4m3_bits(np.array([1e9, -1e9], np.float32)).tolist() == [0x7E, 0xFE]


@pytest.mark.skipif(not hasattr(torch, "float8_e4m3fn"), reason="no fp8 in torch")
def test_e4m3_matches_torch_in_range():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(20000) * 50).astype(np.float32)
    x = x[np.abs(x) <= 448]
    ref = torch.from_numpy(x).to(torch.float8_e4m3fn).view(torch.uint8).numpy()
    assert np.array_equal(fp8.f32_to_e4m3_bits(x), ref)


def test_block_scaled_fp8():
    rng = np.random.default_rng(2)
    n, f = 100, 70
    cols = {f"f{i}": (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
            for i in range(f)}
    cols["labels"] = rng.random(n).astype(np.float32)
    lay = L.build_layout([(f"f{i}", L.DT_F32, L.DT_FP8, 1) for i in range(f)]
                         + [("labels", L.DT_F32, L.DT_F32, 1)], fp8_block_scale=True)
    assert lay.scale_offset >= 0
    packed = L.pack_rows(cols, lay)
    payload = packed[:, :f]
    scales = packed[:, lay.scale_offset:lay.scale_offset + 3]
    deq = fp8.dequantize_block_scaled(payload, scales)
    x = np.stack([cols[f"f{i}"] for i in range(f)], axis=1)
    # e4m3 has 3 mantissa bits: relative error <= 2^-4 of the block max
    blockmax = np.repeat(np.abs(np.pad(x, ((0, 0), (0, 26)))).reshape(n, 3, 32).max(2), 32, 1)[:, :f]
    assert np.all(np.abs(deq - x) <= blockmax * 2.0 ** -4 + 1e-30)
    assert np.array_equal(L.unpack_field(packed, lay.field("labels")), cols["labels"])


def test_row_align_pads_pitch_only():
    import pytest
    from ray_shuffling_data_loader_b200.ops import layout as L
    cols = [(f"c{i}", L.DT_I64, L.DT_F32, 1) for i in range(21)]
    base = L.build_layout(cols)
    wide = L.build_layout(cols, row_align=128)
    assert base.row_pitch == 96 and wide.row_pitch == 128
    assert [f.offset for f in base.fields] == [f.offset for f in wide.fields]
    data = {f"c{i}": np.arange(10, dtype=np.int64) * (i + 1) for i in range(21)}
    a, b = L.pack_rows(data, base), L.pack_rows(data, wide)
    assert np.array_equal(a[:, :84], b[:, :84]) and not b[:, 84:].any()
    with pytest.raises(ValueError):
        L.build_layout(cols, row_align=48)
