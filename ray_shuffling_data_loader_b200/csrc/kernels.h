// Host-visible parameter blocks and launchers of the shuffle kernels.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace rsdl {

// Fast path: every source column is a 4-byte (modes 0-2) or 8-byte (modes 3-4)
// scalar (see shuffle_kernels.cu).
// A small scalar field that follows the fast prefix in the packed row (a float32 label
// behind fp8 / bf16 / int64 features, an int64 id behind float features ...). The fast
// kernel loads it straight from its source column and folds it into the 16-byte group
// store of its row, so a mixed row costs no second kernel and no extra sub-sector
// write transaction per row (profiles/README.md round 2, "bytes on the wire").
struct TailField {
  const uint8_t* src;                // source column (chunk-local, like FastParams::cols)
  uint32_t src_code;
  uint32_t dst_code;                 // 4- or 8-byte destination types only
  uint32_t dst_off;                  // byte offset in the row, aligned to the dst itemsize
  uint32_t pad_;
};
#define RSDL_MAX_TAIL_FIELDS 4

struct FastParams {
  alignas(64) CUtensorMap tmap;      // [num_cols][rows] view of the source columns
  uint32_t use_tmap;                 // 0: 1-D bulk copies from `cols` pointers
  PermKeyDev key;
  PlanDev plan;
  const uint8_t* const* cols;        // device array [num_cols] of column bases
  const uint8_t* kinds;              // mode 4: device array [num_cols] of conversion kinds
  uint32_t num_cols;
  uint32_t num_panels;               // ceil(num_cols / panel width)
  unsigned long long n_local;        // rows owned by this rank
  unsigned long long global_offset;  // global index of local row 0
  uint32_t row_pitch;                // destination row pitch (bytes)
  uint32_t scale_offset;             // fp8 mode: byte offset of the UE8M0 scales
  uint32_t sched;                    // producer schedule: 0 loader warps, 1 cooperative
  uint32_t write_end;                // bytes [num_cols*dsz, write_end) of a row are zero-filled
  // bytes [tail_lo, tail_hi) (16-byte multiples, tail_lo >= prefix end) are written by the
  // tail step: zeros plus the tail fields that live there; empty when tail_hi <= tail_lo
  uint32_t tail_lo;
  uint32_t tail_hi;
  uint32_t num_tail;
  uint32_t pad_;
  TailField tail[RSDL_MAX_TAIL_FIELDS];
  uint8_t* dst[RSDL_MAX_TRAINERS];   // epoch-slot base per trainer (local/peer)
};

struct FieldDev {
  const uint8_t* src;
  uint32_t src_code;
  uint32_t dst_code;
  uint32_t dst_off;
  uint32_t width;
};

struct GenericParams {
  PermKeyDev key;
  PlanDev plan;
  const FieldDev* fields;            // device array
  uint32_t num_fields;
  uint32_t rows_per_block;           // filled by the launcher
  unsigned long long n_local;
  unsigned long long global_offset;
  uint32_t row_pitch;
  uint32_t write_lo;                 // byte range of the row this launch owns
  uint32_t write_hi;                 //   (multiples of 4; whole row by default)
  uint8_t* dst[RSDL_MAX_TRAINERS];
};

// Wide path: one list-valued column ([n, width] row-major) -> one field per row.
struct WideParams {
  PermKeyDev key;
  PlanDev plan;
  const uint8_t* src;
  uint32_t width;
  uint32_t src_code;
  uint32_t dst_code;
  uint32_t dst_off;
  unsigned long long n_local;
  unsigned long long global_offset;
  uint32_t row_pitch;
  uint8_t* dst[RSDL_MAX_TRAINERS];
};

struct FlagTargets {
  uint32_t* ptr[RSDL_MAX_TRAINERS];
  uint32_t count;
};

int fast_panel_cols(int mode);
int fast_ctas_per_sm(int mode);
int fast_tile_rows(int mode);
int fast_src_itemsize(int mode);
int fast_max_tile_rows();

void launch_scatter_fast(const FastParams& p, int mode, int grid, cudaStream_t stream);
void launch_scatter_generic(GenericParams p, int grid, cudaStream_t stream);
void launch_scatter_wide(const WideParams& p, int grid, cudaStream_t stream);
void launch_perm_positions(const PermKeyDev& key, const PlanDev& plan,
                           unsigned long long global_offset, unsigned long long n_local,
                           int32_t* trainer, long long* slot, cudaStream_t stream);
void launch_place_rows(const uint8_t* rows, const long long* src_idx, const long long* dst_off,
                       unsigned long long n, uint32_t pitch, uint8_t* dst, cudaStream_t stream);
void launch_key_checksum(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                         uint32_t key_off, unsigned long long* out, cudaStream_t stream);
void launch_batch_sum_f32(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                          uint32_t off, double* out, cudaStream_t stream);
void launch_batch_sum_all_f32(const uint8_t* packed, unsigned long long nbytes, double* out,
                              cudaStream_t stream);
void launch_bulk_store_probe(uint8_t* dst, unsigned long long slots, uint32_t row_bytes,
                             unsigned long long total_rows, int grid, cudaStream_t stream);
void launch_signal_flags(const FlagTargets& t, uint32_t value, cudaStream_t stream);
void launch_wait_flags(const uint32_t* flags, uint32_t count, uint32_t value,
                       unsigned long long timeout_ns, uint32_t* error, cudaStream_t stream);

}  // namespace rsdl
