"""Block-scaled fp8 (MXFP8-style) epilogue - numpy golden.

BASELINE.json's north star asks for a "block-scaled fp8 cast where the
downstream model trains in fp8": the scatter kernel can emit features as
e4m3 with one UE8M0 (power-of-two) scale per 32 consecutive elements, which is
the operand format ``tcgen05.mma.kind::mxf8f6f4.block_scale`` consumes and cuts
the NVLink bytes of the row exchange 4x versus fp32.

Row format when ``RowLayout.scale_offset >= 0``:

    [e4m3 payload: one byte per element][...other fields...][scales: 1 byte / 32 elems][pad]

``scale byte = e + 127`` with ``e`` the smallest exponent such that
``amax_block * 2**-e <= 448`` (see ``block_scale_exponent``), elements are
``x * 2**-e`` rounded to nearest-even and saturated to +-448. ``csrc/shuffle_kernels.cu`` implements the same rule.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

E4M3_MAX = 448.0
BLOCK = 32


def f32_to_e4m3_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> e4m3fn bit patterns, RNE, saturate-to-finite (``__NV_SATFINITE``)."""
    x = np.asarray(x, dtype=np.float32)
    sign = (np.signbit(x)).astype(np.uint8) << 7
    a = np.abs(x).astype(np.float64)
    nan = np.isnan(a)
    a = np.where(nan, 0.0, np.minimum(a, E4M3_MAX))
    out = np.zeros(a.shape, dtype=np.uint8)
    with np.errstate(divide="ignore"):
        e = np.floor(np.log2(np.where(a > 0, a, 1.0))).astype(np.int64)
    normal = a >= 2.0 ** -6
    # normal numbers: step 2**(e-3)
    e_n = np.where(normal, e, 0)
    q = np.rint(a / np.exp2((e_n - 3).astype(np.float64))).astype(np.int64)
    carry = q == 16
    e_n = np.where(carry, e_n + 1, e_n)
    q = np.where(carry, 8, q)
    bits_n = ((e_n + 7) << 3) | (q - 8)
    bits_n = np.minimum(bits_n, 0x7E)  # 448 is the largest finite value
    # subnormals: step 2**-9
    bits_s = np.rint(a / 2.0 ** -9).astype(np.int64)
    out = np.where(normal, bits_n, bits_s).astype(np.uint8)
    out = np.where(nan, np.uint8(0x7F), out)
    return (out | sign).astype(np.uint8)


def e4m3_bits_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8)
    sign = np.where(b & 0x80, -1.0, 1.0)
    exp = ((b >> 3) & 0xF).astype(np.int64)
    man = (b & 0x7).astype(np.float64)
    val = np.where(exp == 0, man * 2.0 ** -9, (8 + man) * np.exp2((exp - 10).astype(np.float64)))
    val = np.where((b & 0x7F) == 0x7F, np.nan, val)
    return (sign * val).astype(np.float32)


def block_scale_exponent(amax: np.ndarray) -> np.ndarray:
    """Shared exponent ``e`` per block (int64): the smallest power of two with
    ``amax * 2**-e <= 448`` (round-up rule, so the block maximum never clips).
    Computed from the fp32 bits of ``amax * (1/448)`` so that host and device
    agree bit for bit: ``e = exponent + (mantissa != 0)``, clamped to
    [-126, 127]; an all-zero block stores scale byte 0 (``e = -127``)."""
    amax = np.asarray(amax, dtype=np.float32)
    v = (amax * np.float32(1.0 / 448.0)).astype(np.float32)
    bits = v.view(np.uint32).astype(np.int64)
    exp_field = (bits >> 23) & 0xFF
    mant = bits & 0x7FFFFF
    e = exp_field - 127 + (mant != 0)
    e = np.clip(e, -126, 127)
    return np.where(amax > 0, e, -127)


def pack_fp8_block_scaled(columns: Dict[str, np.ndarray], layout,
                          row_idx: Optional[np.ndarray], out: np.ndarray) -> None:
    """Fill the fp8 fields and the scale bytes of ``out`` (``[n, pitch]``)."""
    from ray_shuffling_data_loader_b200.ops.layout import DT_FP8
    fp8_fields = [f for f in layout.fields if f.dst_code == DT_FP8]
    if not fp8_fields:
        return
    n = out.shape[0]
    mats = []
    for f in fp8_fields:
        col = columns[f.name]
        vals = col if row_idx is None else col[row_idx]
        mats.append(np.asarray(vals, dtype=np.float32).reshape(n, f.width))
    x = np.concatenate(mats, axis=1)              # [n, E] in field order
    nelem = x.shape[1]
    nblk = (nelem + BLOCK - 1) // BLOCK
    pad = nblk * BLOCK - nelem
    xp = np.pad(x, ((0, 0), (0, pad))) if pad else x
    blocks = xp.reshape(n, nblk, BLOCK)
    amax = np.fmax.reduce(np.abs(blocks), axis=2)      # NaNs never win (device rule)
    amax = np.where(np.isnan(amax), 0.0, amax)
    e = block_scale_exponent(amax)                # [n, nblk]
    scaled = blocks.astype(np.float64) * np.exp2(-e.astype(np.float64))[:, :, None]
    bits = f32_to_e4m3_bits(scaled.astype(np.float32)).reshape(n, nblk * BLOCK)[:, :nelem]
    col0 = 0
    for f in fp8_fields:
        out[:, f.offset:f.offset + f.width] = bits[:, col0:col0 + f.width]
        col0 += f.width
    out[:, layout.scale_offset:layout.scale_offset + nblk] = (e + 127).astype(np.uint8)


def dequantize_block_scaled(payload: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """``payload``: uint8 ``[n, E]`` e4m3 bits, ``scales``: uint8 ``[n, ceil(E/32)]``."""
    n, nelem = payload.shape
    vals = e4m3_bits_to_f32(payload).astype(np.float64)
    e = scales.astype(np.int64) - 127
    rep = np.repeat(e, BLOCK, axis=1)[:, :nelem]
    return (vals * np.exp2(rep.astype(np.float64))).astype(np.float32)
