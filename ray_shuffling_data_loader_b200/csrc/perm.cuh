// K1 perm_index: stateless keyed bijection, bit-identical to ops/perm.py.
//
// Replaces the reference's unseeded np.random.randint reducer assignment
// (shuffle.py:156) and the reducer-side DataFrame.sample(frac=1)
// (shuffle.py:194): one Feistel evaluation per row yields the row's final
// (trainer, slot), so no index array is ever materialised.
#pragma once

#include "common.cuh"

__host__ __device__ __forceinline__ uint32_t rsdl_round_fn(uint32_t x, uint32_t k) {
  uint32_t h = x * 0x9E3779B1u + k;   // wraps mod 2^32 (same as numpy uint32)
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ unsigned long long
rsdl_feistel(unsigned long long x, const PermKeyDev& key) {
  uint32_t left = static_cast<uint32_t>(x >> key.bits_r);
  uint32_t right = static_cast<uint32_t>(x) & key.mask_r;
#pragma unroll
  for (int i = 0; i < RSDL_PERM_ROUNDS; i += 2) {
    left ^= rsdl_round_fn(right, key.k[i]) & key.mask_l;
    right ^= rsdl_round_fn(left, key.k[i + 1]) & key.mask_r;
  }
  return (static_cast<unsigned long long>(left) << key.bits_r) | right;
}

// pi_e(x): cycle-walk until the value is back inside [0, n). 2^bits < 2n, so
// the expected number of evaluations is < 2.
__host__ __device__ __forceinline__ unsigned long long
rsdl_permute(unsigned long long x, const PermKeyDev& key) {
  if (key.n <= 1) return x;
  do {
    x = rsdl_feistel(x, key);
  } while (x >= key.n);
  return x;
}

__host__ __device__ __forceinline__ unsigned long long
rsdl_feistel_inv(unsigned long long x, const PermKeyDev& key) {
  uint32_t left = static_cast<uint32_t>(x >> key.bits_r);
  uint32_t right = static_cast<uint32_t>(x) & key.mask_r;
#pragma unroll
  for (int i = RSDL_PERM_ROUNDS - 2; i >= 0; i -= 2) {
    right ^= rsdl_round_fn(left, key.k[i + 1]) & key.mask_r;
    left ^= rsdl_round_fn(right, key.k[i]) & key.mask_l;
  }
  return (static_cast<unsigned long long>(left) << key.bits_r) | right;
}

__host__ __device__ __forceinline__ unsigned long long
rsdl_permute_inv(unsigned long long y, const PermKeyDev& key) {
  if (key.n <= 1) return y;
  do {
    y = rsdl_feistel_inv(y, key);
  } while (y >= key.n);
  return y;
}

// Global position -> (trainer, slot in the trainer's epoch buffer).
__host__ __device__ __forceinline__ void
rsdl_position_to_dest(unsigned long long pos, const PlanDev& plan,
                      uint32_t* trainer, unsigned long long* slot) {
  if (pos < plan.big) {
    if (plan.q + 1 <= 0xFFFFFFFFull && pos <= 0xFFFFFFFFull) {
      uint32_t d = static_cast<uint32_t>(plan.q + 1);
      uint32_t t = static_cast<uint32_t>(pos) / d;
      *trainer = t;
      *slot = static_cast<uint32_t>(pos) - t * d;
    } else {
      unsigned long long t = pos / (plan.q + 1);
      *trainer = static_cast<uint32_t>(t);
      *slot = pos - t * (plan.q + 1);
    }
  } else {
    unsigned long long rest = pos - plan.big;
    if (plan.q <= 0xFFFFFFFFull && rest <= 0xFFFFFFFFull) {
      uint32_t d = static_cast<uint32_t>(plan.q);
      uint32_t t = static_cast<uint32_t>(rest) / d;
      *trainer = plan.rem + t;
      *slot = static_cast<uint32_t>(rest) - t * d;
    } else {
      unsigned long long t = rest / plan.q;
      *trainer = plan.rem + static_cast<uint32_t>(t);
      *slot = rest - t * plan.q;
    }
  }
}
