"""Packed-row layout: the gather/cast-to-tensor contract (kernel K8).

The reference converts each DataFrame batch column by column with
``torch.as_tensor(column, dtype)`` + ``.view(-1, *shape)`` (reference
``torch_dataset.py:204-236``). Here the cast happens once, inside the shuffle
kernel's epilogue: every shuffled row is written as one packed record

    [field0 | field1 | ... | pad]            (``row_pitch`` bytes, 16 B multiple)

where a field is one source column cast to its destination dtype (``width``
elements for list-valued columns). A batch is then a ``[B, row_pitch]`` byte
matrix and each feature tensor is a strided zero-copy view of it, so the
``List[(B,1) Tensor], (B,1) label`` contract is met without any per-batch work.

dtype codes are shared with ``csrc/common.cuh``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# code -> (name, itemsize)
DT_U8, DT_I8, DT_I16, DT_I32, DT_I64, DT_F16, DT_BF16, DT_F32, DT_F64, DT_FP8 = range(10)
DT_BOOL = 10
_ITEMSIZE = {DT_U8: 1, DT_I8: 1, DT_I16: 2, DT_I32: 4, DT_I64: 8, DT_F16: 2,
             DT_BF16: 2, DT_F32: 4, DT_F64: 8, DT_FP8: 1, DT_BOOL: 1}
_NAMES = {DT_U8: "uint8", DT_I8: "int8", DT_I16: "int16", DT_I32: "int32",
          DT_I64: "int64", DT_F16: "float16", DT_BF16: "bfloat16",
          DT_F32: "float32", DT_F64: "float64", DT_FP8: "float8_e4m3fn",
          DT_BOOL: "bool"}
_NP_TO_CODE = {np.dtype(np.uint8): DT_U8, np.dtype(np.int8): DT_I8,
               np.dtype(np.int16): DT_I16, np.dtype(np.int32): DT_I32,
               np.dtype(np.int64): DT_I64, np.dtype(np.float16): DT_F16,
               np.dtype(np.float32): DT_F32, np.dtype(np.float64): DT_F64,
               np.dtype(np.bool_): DT_BOOL}
_CODE_TO_NP = {DT_U8: np.uint8, DT_I8: np.int8, DT_I16: np.int16,
               DT_I32: np.int32, DT_I64: np.int64, DT_F16: np.float16,
               DT_BF16: np.uint16, DT_F32: np.float32, DT_F64: np.float64,
               DT_FP8: np.uint8, DT_BOOL: np.bool_}


def itemsize(code: int) -> int:
    return _ITEMSIZE[code]


def dtype_name(code: int) -> str:
    return _NAMES[code]


def code_from_numpy(dtype) -> int:
    dt = np.dtype(dtype)
    if dt not in _NP_TO_CODE:
        raise TypeError(f"unsupported column dtype {dt}")
    return _NP_TO_CODE[dt]


def numpy_storage_dtype(code: int):
    """numpy dtype used to *store* a code (bf16/fp8 are stored as raw bits)."""
    return _CODE_TO_NP[code]


def code_from_torch(dtype) -> int:
    import torch
    table = {torch.uint8: DT_U8, torch.int8: DT_I8, torch.int16: DT_I16,
             torch.int32: DT_I32, torch.int64: DT_I64, torch.float16: DT_F16,
             torch.bfloat16: DT_BF16, torch.float32: DT_F32,
             torch.float64: DT_F64, torch.bool: DT_BOOL}
    if hasattr(torch, "float8_e4m3fn"):
        table[torch.float8_e4m3fn] = DT_FP8
    if dtype not in table:
        raise TypeError(f"unsupported torch dtype {dtype}")
    return table[dtype]


def torch_dtype(code: int):
    import torch
    table = {DT_U8: torch.uint8, DT_I8: torch.int8, DT_I16: torch.int16,
             DT_I32: torch.int32, DT_I64: torch.int64, DT_F16: torch.float16,
             DT_BF16: torch.bfloat16, DT_F32: torch.float32,
             DT_F64: torch.float64, DT_BOOL: torch.bool}
    if hasattr(torch, "float8_e4m3fn"):
        table[DT_FP8] = torch.float8_e4m3fn
    return table[code]


@dataclass(frozen=True)
class Field:
    name: str          # source column name
    src_code: int      # dtype code of the source column
    dst_code: int      # dtype code inside the packed row
    offset: int        # byte offset inside the row
    width: int = 1     # elements per row (list-valued columns have width > 1)

    @property
    def dst_bytes(self) -> int:
        return self.width * itemsize(self.dst_code)

    @property
    def src_bytes(self) -> int:
        return self.width * itemsize(self.src_code)


@dataclass(frozen=True)
class RowLayout:
    fields: Tuple[Field, ...]
    row_pitch: int
    # Optional block-scaled fp8 section (see ops/fp8.py): when set, fields with
    # dst_code == DT_FP8 are scaled per 32-element block by an UE8M0 exponent
    # stored at ``scale_offset + block_index``.
    scale_offset: int = -1
    # User-facing column order as indices into ``fields`` (None: storage order).
    # ``dataframe_layout`` may store columns in a kernel-friendly order while
    # DataFrames / ``DeviceBatch.columns`` keep the file's order.
    display_order: Optional[Tuple[int, ...]] = None

    def field(self, name: str) -> Field:
        for f in self.fields:
            if f.name == name:
                return f
        raise KeyError(name)

    @property
    def names(self) -> List[str]:
        """Column names in user-facing order."""
        return [f.name for f in self.display_fields]

    @property
    def display_fields(self) -> Tuple[Field, ...]:
        if self.display_order is None:
            return self.fields
        return tuple(self.fields[i] for i in self.display_order)

    @property
    def payload_bytes(self) -> int:
        return sum(f.dst_bytes for f in self.fields)

    @property
    def uniform_code(self) -> Optional[int]:
        """dst dtype code if every field shares one dst dtype and the fields are
        densely packed from offset 0 (=> a ``[B, F]`` matrix view exists)."""
        if not self.fields:
            return None
        code = self.fields[0].dst_code
        off = 0
        for f in self.fields:
            if f.dst_code != code or f.offset != off:
                return None
            off += f.dst_bytes
        return code

    @property
    def is_fast_path(self) -> bool:
        """True when the TMA fast kernel applies: every source column is a
        4-byte scalar copied bit-for-bit (or f32 -> bf16/fp8), dense from 0."""
        code = self.uniform_code
        if code is None:
            return False
        srcs = {f.src_code for f in self.fields}
        if any(f.width != 1 for f in self.fields) or len(srcs) != 1:
            return False
        src = next(iter(srcs))
        if itemsize(src) != 4:
            return False
        if code == src:
            return True
        return src == DT_F32 and code in (DT_BF16, DT_FP8)


def _align(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def build_layout(columns: Sequence[Tuple[str, int, int, int]],
                 fp8_block_scale: bool = False, row_align: int = 0) -> RowLayout:
    """``columns``: (name, src_code, dst_code, width) in output order. Every
    field is aligned to its destination itemsize; the pitch to 16 bytes.

    ``row_align`` (opt-in, a power of two >= 32) rounds the pitch up further.
    Rows that are a multiple of the 128-byte L2 line scatter at 80-91 % of the HBM
    copy peak while 64-192-byte rows reach 49-65 % (profiles/README.md, "row
    size"), so padding e.g. 84-byte rows to 128 trades 33 % more bytes for full
    lines; it is not the default because it also grows every epoch slot."""
    fields = []
    off = 0
    for name, src, dst, width in columns:
        off = _align(off, itemsize(dst))
        fields.append(Field(name, src, dst, off, width))
        off += width * itemsize(dst)
    scale_offset = -1
    if fp8_block_scale:
        nelem = sum(f.width for f in fields if f.dst_code == DT_FP8)
        # 16-aligned so the kernel's 16-byte payload stores never touch a scale
        scale_offset = _align(off, 16)
        off = scale_offset + (nelem + 31) // 32
    # Rows are scattered one at a time: keep every row on its own 32-byte DRAM
    # sectors (no read-modify-write of a sector shared with a neighbour row).
    pitch = _align(off, 16) if off <= 16 else _align(off, 32)
    if row_align:
        if row_align < 32 or row_align & (row_align - 1):
            raise ValueError("row_align must be a power of two >= 32")
        pitch = _align(pitch, row_align)
    return RowLayout(tuple(fields), max(16, pitch), scale_offset)


# (8-byte source, 4-byte destination) casts the TMA kernel converts in its epilogue
# (csrc/shuffle_kernels.cu mode 4); runtime/device_engine.py maps them to kinds.
TMA_CASTS_8_TO_4 = ((DT_I64, DT_F32), (DT_F64, DT_F32), (DT_I64, DT_I32))


def tma_class(src_code: int, dst_code: int, width: int = 1):
    """Key of the scatter_tma_kernel mode a scalar column can ride in (columns
    with equal keys can share one launch as a dense run), or ``None``."""
    if width != 1:
        return None
    ssz, dsz = itemsize(src_code), itemsize(dst_code)
    if ssz == 4:
        if dst_code == src_code or (src_code == DT_F32 and dst_code == DT_BF16):
            return (4, src_code, dst_code)
        return None
    if ssz == 8:
        if dsz == 8 and dst_code == src_code:
            return (8, 8)
        if dsz == 4 and (src_code, dst_code) in TMA_CASTS_8_TO_4:
            return (8, 4)
    return None


def tma_friendly_order(columns: Sequence[Tuple[str, int, int, int]]) -> List[int]:
    """Indices of ``columns`` ((name, src, dst, width)) with the largest
    TMA-eligible class (by bytes, >= 4 columns) moved to the front; relative
    order is otherwise preserved. Identity when nothing is gained."""
    classes: Dict[tuple, List[int]] = {}
    for i, (_, src, dst, width) in enumerate(columns):
        key = tma_class(src, dst, width)
        if key is not None:
            classes.setdefault(key, []).append(i)
    best = max(classes.values(),
               key=lambda idx: (len(idx) * itemsize(columns[idx[0]][2]), -idx[0]), default=[])
    if len(best) < 4 or best == list(range(len(best))):
        return list(range(len(columns)))
    chosen = set(best)
    return best + [i for i in range(len(columns)) if i not in chosen]


def dataframe_layout(schema: Dict[str, Tuple[int, int]], row_align: int = 0,
                     optimize: bool = True) -> RowLayout:
    """All columns, native dtypes (``schema``: name -> (code, width)). This is
    what plain ``ShufflingDataset`` shuffles (whole rows, like the reference's
    DataFrames).

    *Storage* order: the TMA scatter kernel takes a dense prefix of same-class
    scalar columns (identical 4-byte dtype, or any mix of 8-byte dtypes), so with
    ``optimize`` the largest such class is stored first and everything else behind
    it - an int64 ``key`` in front of 64 float32 features would otherwise push
    the whole row onto the generic kernel. *User-facing* order (DataFrame columns,
    ``DeviceBatch.columns``) stays the file's: see ``RowLayout.display_order``."""
    items = list(schema.items())
    order = list(range(len(items)))
    if optimize and items:
        order = tma_friendly_order([(n, c, c, w) for n, (c, w) in items])
    lay = build_layout([(items[i][0], items[i][1][0], items[i][1][0], items[i][1][1])
                        for i in order], row_align=row_align)
    if order == list(range(len(items))):
        return lay
    inverse = [0] * len(order)
    for stored, original in enumerate(order):
        inverse[original] = stored
    return RowLayout(lay.fields, lay.row_pitch, lay.scale_offset, tuple(inverse))


# ---------------------------------------------------------------------------
# numpy golden for the cast + pack epilogue
# ---------------------------------------------------------------------------

def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns
    (matches ``__float2bfloat16_rn``; NaN is canonicalised to 0x7FFF like CUDA)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32).astype(np.uint64)
    rounded = (bits + np.uint64(0x7FFF) + ((bits >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)
    out = rounded.astype(np.uint16)
    out[np.isnan(x)] = 0x7FFF
    return out


def cast_column(values: np.ndarray, src_code: int, dst_code: int) -> np.ndarray:
    """Cast with the device kernel's rules; result uses numpy_storage_dtype."""
    if dst_code == src_code:
        return values
    if dst_code == DT_BF16:
        return f32_to_bf16_bits(values.astype(np.float32))
    if dst_code == DT_FP8:
        from ray_shuffling_data_loader_b200.ops.fp8 import f32_to_e4m3_bits
        return f32_to_e4m3_bits(values.astype(np.float32))
    if dst_code == DT_BOOL:
        return values != 0
    with np.errstate(all="ignore"):
        return values.astype(_CODE_TO_NP[dst_code])


def pack_rows(columns: Dict[str, np.ndarray], layout: RowLayout,
              row_idx: Optional[np.ndarray] = None) -> np.ndarray:
    """Gather ``row_idx`` rows of the columnar table and pack them:
    ``uint8[len(row_idx), row_pitch]``. ``row_idx=None`` packs all rows in order."""
    any_col = columns[layout.fields[0].name]
    n = len(any_col) if row_idx is None else len(row_idx)
    out = np.zeros((n, layout.row_pitch), dtype=np.uint8)
    if layout.scale_offset >= 0:
        from ray_shuffling_data_loader_b200.ops.fp8 import pack_fp8_block_scaled
        pack_fp8_block_scaled(columns, layout, row_idx, out)
        skip = {f.name for f in layout.fields if f.dst_code == DT_FP8}
    else:
        skip = set()
    for f in layout.fields:
        if f.name in skip:
            continue
        col = columns[f.name]
        vals = col if row_idx is None else col[row_idx]
        vals = cast_column(vals, f.src_code, f.dst_code)
        vals = np.ascontiguousarray(vals).reshape(n, f.width)
        raw = vals.view(np.uint8).reshape(n, f.dst_bytes)
        out[:, f.offset:f.offset + f.dst_bytes] = raw
    return out


def unpack_field(packed: np.ndarray, f: Field) -> np.ndarray:
    """Inverse of pack for one field: returns ``[n]`` or ``[n, width]`` in the
    storage dtype (a copy)."""
    n = packed.shape[0]
    # always a private copy (ascontiguousarray would alias the epoch buffer when
    # one field fills the whole row pitch, and that buffer is recycled)
    raw = np.array(packed[:, f.offset:f.offset + f.dst_bytes], order="C", copy=True)
    vals = raw.view(_CODE_TO_NP[f.dst_code]).reshape(n, f.width)
    return vals[:, 0] if f.width == 1 else vals
