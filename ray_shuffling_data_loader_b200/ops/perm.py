"""Stateless bijective row permutation (kernel K1 ``perm_index``) - numpy golden.

The reference shuffles with two *unseeded* numpy draws per epoch
(``np.random.randint`` in the mapper, ``DataFrame.sample(frac=1)`` in the
reducer; reference ``ray_shuffling_data_loader/shuffle.py:156,194``), which
makes an epoch irreproducible and gives trainers unequal row counts.

Here one epoch's shuffle is a single keyed bijection ``pi_e: [0, N) -> [0, N)``
evaluated per row, never materialised:

* an alternating (unbalanced-safe) Feistel network over ``2**bits >= N`` with
  cycle walking back into ``[0, N)``;
* round keys derived from ``(seed, epoch)`` by splitmix64, so every rank and the
  CUDA kernel (``csrc/perm.cuh``) compute exactly the same mapping with no
  communication;
* invertible (``inverse``), so a destination can enumerate its sources (used by
  the pull-mode gather and by the exactly-once tests).

Everything in this file is mirrored bit-for-bit by ``csrc/perm.cuh``; the GPU
tests diff the two.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

NUM_ROUNDS = 6
_GOLDEN = 0x9E3779B97F4A7C15
_U64 = (1 << 64) - 1


def _splitmix64(state: int) -> Tuple[int, int]:
    """One splitmix64 step on python ints -> (new_state, output)."""
    state = (state + _GOLDEN) & _U64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _U64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _U64
    z = z ^ (z >> 31)
    return state, z


@dataclass(frozen=True)
class PermKey:
    """Everything a kernel needs to evaluate ``pi_e``: passed by value."""
    n: int                 # domain size N
    bits_l: int            # width of the left Feistel half
    bits_r: int            # width of the right Feistel half
    keys: Tuple[int, ...]  # NUM_ROUNDS 32-bit round keys

    @property
    def mask_l(self) -> int:
        return (1 << self.bits_l) - 1

    @property
    def mask_r(self) -> int:
        return (1 << self.bits_r) - 1

    def as_words(self) -> Tuple[int, ...]:
        """Flat tuple for the native call: (n, bits_l, bits_r, k0..k5)."""
        return (self.n, self.bits_l, self.bits_r) + tuple(self.keys)


def make_key(n: int, seed: int, epoch: int) -> PermKey:
    """Derive the epoch's permutation key from ``(seed, epoch)``."""
    if n < 0:
        raise ValueError("n must be non-negative")
    bits = max(2, int(n - 1).bit_length()) if n > 1 else 2
    bits_l = bits // 2
    bits_r = bits - bits_l
    state = (int(seed) & _U64) ^ ((int(epoch) * 0xD1B54A32D192ED03) & _U64)
    keys = []
    for _ in range(NUM_ROUNDS):
        state, out = _splitmix64(state)
        keys.append(out & 0xFFFFFFFF)
    return PermKey(int(n), bits_l, bits_r, tuple(keys))


_BLOCK = 1 << 16   # rows per vectorised block (keeps temporaries in L2)


def _round_fn(x: np.ndarray, k: int) -> np.ndarray:
    """32-bit round function on uint32 lanes (murmur3 fmix32 of x*phi + k);
    uint32 arithmetic wraps, exactly like the device code."""
    h = x * np.uint32(0x9E3779B1)
    h += np.uint32(k)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def _feistel(x: np.ndarray, key: PermKey, inverse: bool = False) -> np.ndarray:
    mask_l = np.uint32(key.mask_l)
    mask_r = np.uint32(key.mask_r)
    br = np.uint64(key.bits_r)
    left = (x >> br).astype(np.uint32)
    right = (x & np.uint64(key.mask_r)).astype(np.uint32)
    rounds = range(NUM_ROUNDS - 1, -1, -1) if inverse else range(NUM_ROUNDS)
    for i in rounds:
        if i % 2 == 0:
            left ^= _round_fn(right, key.keys[i]) & mask_l
        else:
            right ^= _round_fn(left, key.keys[i]) & mask_r
    return (left.astype(np.uint64) << br) | right.astype(np.uint64)


def _walk_block(x: np.ndarray, key: PermKey, inverse: bool) -> np.ndarray:
    out = _feistel(x, key, inverse)
    n = np.uint64(key.n)
    bad = np.nonzero(out >= n)[0]
    # Cycle walking: re-apply until the value falls back inside [0, N).
    while bad.size:
        out[bad] = _feistel(out[bad], key, inverse)
        bad = bad[out[bad] >= n]
    return out


def _walk(x: np.ndarray, key: PermKey, inverse: bool) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    if key.n <= 1:
        return x.copy()
    if key.bits_l > 32 or key.bits_r > 32:
        raise ValueError("domains above 2**64 rows are not supported")
    flat = x.reshape(-1)
    out = np.empty_like(flat)
    with np.errstate(over="ignore"):
        for s in range(0, flat.size, _BLOCK):
            out[s:s + _BLOCK] = _walk_block(flat[s:s + _BLOCK], key, inverse)
    return out.reshape(x.shape)


def permute(x, key: PermKey) -> np.ndarray:
    """``pi_e(x)`` for an array of global row indices (uint64 in, uint64 out)."""
    return _walk(x, key, inverse=False)


def inverse(y, key: PermKey) -> np.ndarray:
    """``pi_e^{-1}(y)``: which source row lands at global position ``y``."""
    return _walk(y, key, inverse=True)


def full_permutation(n: int, seed: int, epoch: int) -> np.ndarray:
    """Materialised permutation (testing / tiny inputs only)."""
    return permute(np.arange(n, dtype=np.uint64), make_key(n, seed, epoch))
