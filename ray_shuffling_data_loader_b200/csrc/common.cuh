// Shared definitions for the sm_100a shuffle kernels and the host runtime.
//
// The reference has no native code at all: every native role it
// leans on (task scheduler, plasma object store, object transfer, Arrow decode)
// lives in Ray / Arrow. This directory is the B200-native replacement for those
// roles: HBM arenas + CUDA-IPC peer mapping, the epoch ring's signal words, and
// the fused permutation/gather/cast/pack/scatter kernels.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#define RSDL_MAX_TRAINERS 64
#define RSDL_PERM_ROUNDS 6

// dtype codes - keep in sync with ops/layout.py
enum : uint32_t {
  DT_U8 = 0, DT_I8 = 1, DT_I16 = 2, DT_I32 = 3, DT_I64 = 4, DT_F16 = 5,
  DT_BF16 = 6, DT_F32 = 7, DT_F64 = 8, DT_FP8 = 9, DT_BOOL = 10
};

__host__ __device__ inline uint32_t rsdl_itemsize(uint32_t code) {
  switch (code) {
    case DT_I16: case DT_F16: case DT_BF16: return 2;
    case DT_I32: case DT_F32: return 4;
    case DT_I64: case DT_F64: return 8;
    default: return 1;
  }
}

// Keyed bijection pi_e over [0, n): mirrors ops/perm.py::PermKey.
struct PermKeyDev {
  unsigned long long n;
  uint32_t bits_r;
  uint32_t mask_l;
  uint32_t mask_r;
  uint32_t k[RSDL_PERM_ROUNDS];
};

// Balanced split of the permuted position space over trainers: mirrors
// ops/plan.py::ShufflePlan.position_to_trainer.
struct PlanDev {
  unsigned long long q;     // rows per trainer (floor)
  unsigned long long big;   // rem * (q + 1)
  uint32_t rem;             // first `rem` trainers own q + 1 rows
  uint32_t num_trainers;
  // Destination-chunk pass filter (K7): a launch only delivers rows whose slot in
  // the trainer's epoch buffer lies in [slot_lo, slot_hi); a full epoch is
  // [0, ~0). See DeviceShuffleEngine.chunk_passes.
  unsigned long long slot_lo;
  unsigned long long slot_hi;
};

#define RSDL_CUDA_CHECK(expr)                                                   \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      throw std::runtime_error(std::string(#expr) + " failed: " +               \
                               cudaGetErrorString(_e));                         \
    }                                                                           \
  } while (0)
