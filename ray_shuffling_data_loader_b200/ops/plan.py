"""Epoch shuffle plan: who owns which global positions, chunking and batching.

The reference assigns reducer outputs to trainers contiguously with
``np.array_split(shuffled, num_trainers)`` (reference ``shuffle.py:125``) and
re-batches variable-sized reducer outputs on the trainer
(``dataset.py:144-168``). Because reducer sizes are multinomial there, trainers
get unequal row counts. Here the permuted position space ``[0, N)`` is split
into ``num_trainers`` contiguous ranges balanced to +-1 row, each trainer range
is split into its share of the ``num_reducers`` "reducer chunks" (the unit that
is handed to ``BatchConsumer.consume`` and that carries a completion flag), and
batches are plain ``batch_size`` strides over the trainer range - so every
batch is contiguous at birth (kernel K6 "batch packing" is fused away).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


def balanced_split(n: int, parts: int) -> List[Tuple[int, int]]:
    """``np.array_split`` boundaries: first ``n % parts`` parts get one extra."""
    q, rem = divmod(n, parts)
    out = []
    start = 0
    for i in range(parts):
        size = q + (1 if i < rem else 0)
        out.append((start, start + size))
        start += size
    return out


@dataclass(frozen=True)
class ShufflePlan:
    """Static (epoch independent) geometry of the shuffle."""
    num_rows: int          # N, global
    num_trainers: int      # T
    num_reducers: int      # R (total, over all trainers)
    batch_size: int
    drop_last: bool = False

    def __post_init__(self):
        if self.num_trainers < 1:
            raise ValueError("num_trainers must be >= 1")
        if self.num_reducers < 1:
            raise ValueError("num_reducers must be >= 1")
        if self.batch_size < 1:
            raise ValueError("batch_size must be >= 1")

    # ---- trainer ranges -------------------------------------------------
    @property
    def rows_q(self) -> int:
        return self.num_rows // self.num_trainers

    @property
    def rows_rem(self) -> int:
        return self.num_rows % self.num_trainers

    def trainer_range(self, t: int) -> Tuple[int, int]:
        q, rem = self.rows_q, self.rows_rem
        start = t * q + min(t, rem)
        return start, start + q + (1 if t < rem else 0)

    def trainer_rows(self, t: int) -> int:
        a, b = self.trainer_range(t)
        return b - a

    @property
    def max_trainer_rows(self) -> int:
        return self.rows_q + (1 if self.rows_rem else 0)

    def position_to_trainer(self, pos: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Vectorised (trainer, slot) of global positions - mirrors the device
        code in ``csrc/perm.cuh::position_to_dest``."""
        pos = np.asarray(pos, dtype=np.uint64)
        q = np.uint64(self.rows_q)
        rem = np.uint64(self.rows_rem)
        big = rem * (q + np.uint64(1))
        in_big = pos < big
        qq = np.uint64(max(int(q), 1))
        t_big = pos // (q + np.uint64(1))
        s_big = pos - t_big * (q + np.uint64(1))
        rest = np.where(in_big, np.uint64(0), pos - big)
        t_small = rem + rest // qq
        s_small = rest - (rest // qq) * qq
        trainer = np.where(in_big, t_big, t_small).astype(np.int64)
        slot = np.where(in_big, s_big, s_small).astype(np.int64)
        return trainer, slot

    # ---- reducer chunks -------------------------------------------------
    def reducers_of_trainer(self, t: int) -> int:
        """How many reducer chunks trainer ``t`` receives: the sizes of
        ``np.array_split(range(R), T)`` (reference ``shuffle.py:125``), but
        never zero so a trainer always has at least one chunk."""
        a, b = balanced_split(self.num_reducers, self.num_trainers)[t]
        return max(1, b - a)

    def trainer_chunks(self, t: int) -> List[Tuple[int, int]]:
        """Row ranges (relative to the trainer's buffer) of its chunks."""
        return balanced_split(self.trainer_rows(t), self.reducers_of_trainer(t))

    # ---- batches --------------------------------------------------------
    def num_batches(self, t: int) -> int:
        rows = self.trainer_rows(t)
        full, tail = divmod(rows, self.batch_size)
        return full + (1 if (tail and not self.drop_last) else 0)

    def batch_range(self, t: int, b: int) -> Tuple[int, int]:
        rows = self.trainer_rows(t)
        start = b * self.batch_size
        return start, min(start + self.batch_size, rows)

    # ---- source ranges --------------------------------------------------
    def source_range(self, rank: int, world: int) -> Tuple[int, int]:
        """Global row range ingested (and scattered) by process ``rank``."""
        return balanced_split(self.num_rows, world)[rank]
