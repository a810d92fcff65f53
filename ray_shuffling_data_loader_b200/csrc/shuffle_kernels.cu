// Fused shuffle kernels for sm_100a (kernels K1, K2, K5, K6, K8, K11, K13 of SURVEY.md 2.4).
//
// What the reference does on CPUs across three Ray stages - R boolean-mask
// partition passes per file (shuffle.py:156-161), the mapper->reducer
// all-to-all through the object store (shuffle.py:120-123), pd.concat +
// sample(frac=1) in the reducer (shuffle.py:192-194), re-batching
// (dataset.py:144-168) and per-column torch.as_tensor (torch_dataset.py:209-235)
// - is one kernel here:
//
//   scatter_tma_kernel    warp-specialised, one CTA per SM: loader warps move a
//                         [cols x TILE_ROWS] tile of the local *columnar* table
//                         into padded shared memory with TMA bulk copies; index
//                         warps evaluate the Feistel bijection for the tile's rows
//                         and publish each row's final address - (trainer, slot)
//                         in local or *peer* HBM (VMM / CUDA-IPC mapped, NVLink 5
//                         / NVSwitch) - through shared memory; consumer warps do a
//                         conflict-free 4x4 register transposition to row-major,
//                         the cast (f32 | bf16 | block-scaled e4m3 | int64/float64
//                         copy | int64/float64 -> f32/int32) and 128 B coalesced
//                         vector stores of every row straight to that address.
//   scatter_wide_kernel   list-valued columns (images): permuted row copy + cast.
//   scatter_generic_kernel everything else (1-2-byte sources, exotic casts),
//                         staged through shared memory.
//
// No NCCL call and no intermediate buffer sit on this path; NCCL all_to_all is
// only the baseline (parallel/nccl_baseline.py).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "kernels.h"
#include "perm.cuh"

// Tile geometry of the f32/bf16 fast path (overridable for experiments).
// Measured on B200 (profiles/kbench_v4.jsonl, kbench_v5.jsonl): 128-row tiles,
// 4 stages, 1 CTA/SM, 1-D bulk loads = 1.08 ms for 3.2 GB in + 3.2 GB out (90% of
// the measured HBM copy peak); 256-row tiles or 2 CTAs/SM are 6-10% slower.
#ifndef RSDL_TILE_F32
#define RSDL_TILE_F32 128
#endif
#ifndef RSDL_STAGES_F32
#define RSDL_STAGES_F32 4
#endif
#ifndef RSDL_CTAS_F32
#define RSDL_CTAS_F32 1
#endif
#ifndef RSDL_PANEL_F32
#define RSDL_PANEL_F32 64
#endif
// Consumer warp groups (8 warps each) that take alternate pipeline stages.
// Measured (profiles/kbench_v8/v10/v20.jsonl): a second group changes nothing
// for any mode or width - the consumers are never the limiter once the row
// padding is zero-filled - so one group is the default; kept as an A/B knob.
#ifndef RSDL_CGROUPS_F32
#define RSDL_CGROUPS_F32 1
#endif
#ifndef RSDL_CGROUPS_64
#define RSDL_CGROUPS_64 1
#endif
#ifndef RSDL_TILE_64
#define RSDL_TILE_64 128
#endif
#ifndef RSDL_STAGES_64
#define RSDL_STAGES_64 4
#endif

namespace rsdl {

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copy (TMA, non-tensor form), vector ld/st
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// RSDL_RACECHECK build (tools/sanitize.sh racecheck): compute-sanitizer's racecheck
// models bar.sync / bar.arrive but not mbarrier arrive / try_wait issued from inline
// PTX, so it reports the (mbarrier-ordered) index -> consumer hand-off of dptr[] as
// a potential RAW hazard on every tile. In this build the hand-off is *also* fenced
// with a named barrier per stage (ids 1..STAGES: the publishing index warp arrives,
// the consumer warps sync), which the tool understands; the production build relies
// on the mbarrier alone. The extra barrier cannot deadlock: it is reached in exactly
// the same order as the mbarrier phases of the same stage.
#ifdef RSDL_RACECHECK
__device__ __forceinline__ void rc_arrive(int stage, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(stage + 1), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void rc_sync(int stage, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(stage + 1), "r"(nthreads) : "memory");
}
#else
__device__ __forceinline__ void rc_arrive(int, int) {}
__device__ __forceinline__ void rc_sync(int, int) {}
#endif

// cp.async.bulk global -> shared, completion reported on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src,
                                            uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// cp.async.bulk.tensor 2-D tile load through a tensor map (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int x, int y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ unsigned long long lds64(const void* p) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ void sts64(void* p, unsigned long long v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}

// 16-byte store to a (possibly peer-mapped) global address.
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#ifdef RSDL_EXPERIMENT_NO_STORE
  // experiment only (never in the shipped build): keep the value dependency,
  // drop the store, to measure the pipeline without its DRAM writes
  if (a == 0x7fc12345u && b == c) asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  return;
#endif
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t f32_to_e4m3(float x) {
  return static_cast<uint32_t>(__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3));
}

// Two values per conversion instruction (cvt.rn.satfinite.e4m3x2.f32, same rounding as
// the scalar form): byte 0 = lo, byte 1 = hi. Halves the F2FP count of the fp8 epilogue
// and drops two of the three shift/or pairs per packed word.
__device__ __forceinline__ uint32_t f32x2_to_e4m3x2(float lo, float hi) {
  return static_cast<uint32_t>(
      __nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3));
}

// Destination address of global source row `gi` (0 when the row does not exist).
__device__ __forceinline__ unsigned long long
dest_pointer(unsigned long long gi, const PermKeyDev& key, const PlanDev& plan,
             uint8_t* const* dst, uint32_t row_pitch) {
  unsigned long long pos = rsdl_permute(gi, key);
  uint32_t trainer;
  unsigned long long slot;
  rsdl_position_to_dest(pos, plan, &trainer, &slot);
  if (slot < plan.slot_lo || slot >= plan.slot_hi) return 0ull;   // another pass delivers it
  return reinterpret_cast<unsigned long long>(dst[trainer]) + slot * row_pitch;
}

// Destination bits (low 4 or 8 bytes) of a tail field whose raw source value is `raw`.
// Deliberately a short list - bit copies of 4- and 8-byte types and the three 8 -> 4
// byte conversions of mode 4 (same rounding: RNE / truncation) - so that the
// non-inlined tail_step stays small; anything else remains with the generic kernel.
__device__ __forceinline__ unsigned long long tail_convert(const TailField& f, unsigned long long raw) {
  if (f.dst_code == f.src_code) return raw;
  if (f.dst_code == DT_I32) return static_cast<uint32_t>(raw);                       // int64 -> int32
  if (f.src_code == DT_F64)                                                          // float64 -> f32
    return __float_as_uint(__double2float_rn(__longlong_as_double(static_cast<long long>(raw))));
  return __float_as_uint(__ll2float_rn(static_cast<long long>(raw)));               // int64 -> f32
}

// Bytes [tail_lo, tail_hi) of 4 consecutive rows (destinations d0..d3, 0 = skip; chunk-
// local source rows row0..row0+3, row0 % 4 == 0): zeros plus the small fields that
// follow the prefix. Lane q of the row group writes the 16-byte groups q, q+8, ... - every
// row's tail is again whole 16-byte stores. A field's values for the 4 rows are one
// 16-byte (4-byte sources) or two 16-byte (8-byte sources) vector loads from its source
// column (columns are padded to whole tiles, so rows past the end are readable).
__device__ __noinline__ void tail_step(const FastParams* p, unsigned long long d0,
                                       unsigned long long d1, unsigned long long d2,
                                       unsigned long long d3, unsigned long long row0, int q) {
  const uint32_t ngroups = (p->tail_hi - p->tail_lo) >> 4;
  for (uint32_t tg = q; tg < ngroups; tg += 8) {
    const uint32_t off = p->tail_lo + (tg << 4);
    uint32_t w[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j][0] = w[j][1] = w[j][2] = w[j][3] = 0u;
    for (uint32_t t = 0; t < p->num_tail; ++t) {
      const uint32_t doff = p->tail[t].dst_off;
      if (doff < off || doff >= off + 16u) continue;
      unsigned long long raw[4];
      if (rsdl_itemsize(p->tail[t].src_code) == 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p->tail[t].src + row0 * 4ull));
        raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w;
      } else {
        const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(p->tail[t].src + row0 * 8ull);
        const ulonglong2 a = __ldg(s2), b = __ldg(s2 + 1);
        raw[0] = a.x; raw[1] = a.y; raw[2] = b.x; raw[3] = b.y;
      }
      const uint32_t wi = (doff - off) >> 2;
      const bool wide = rsdl_itemsize(p->tail[t].dst_code) == 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned long long bits = rsdl_itemsize(p->tail[t].src_code) == 4
            ? raw[j] : tail_convert(p->tail[t], raw[j]);
        const uint32_t lo = static_cast<uint32_t>(bits), hi = static_cast<uint32_t>(bits >> 32);
        if (wi == 0) { w[j][0] = lo; if (wide) w[j][1] = hi; }
        else if (wi == 1) { w[j][1] = lo; }
        else if (wi == 2) { w[j][2] = lo; if (wide) w[j][3] = hi; }
        else { w[j][3] = lo; }
      }
    }
    if (d0) stg128(reinterpret_cast<void*>(d0 + off), w[0][0], w[0][1], w[0][2], w[0][3]);
    if (d1) stg128(reinterpret_cast<void*>(d1 + off), w[1][0], w[1][1], w[1][2], w[1][3]);
    if (d2) stg128(reinterpret_cast<void*>(d2 + off), w[2][0], w[2][1], w[2][2], w[2][3]);
    if (d3) stg128(reinterpret_cast<void*>(d3 + off), w[3][0], w[3][1], w[3][2], w[3][3]);
  }
}

// ---------------------------------------------------------------------------
// K2 shuffle_scatter, fast path: uniform 4-byte source columns
// ---------------------------------------------------------------------------
//
// Tile = TILE_ROWS consecutive local rows x up to PANEL columns. Shared memory
// holds the tile *columnar*: column c at word offset c * P with P = TILE_ROWS+4.
// P/4 is odd, so the 16-byte bank group of element (c, r) is (c + r/4) mod 8.
//
// A consumer lane owns a 4-row x FPL-column block (FPL = 16 / dst itemsize):
// lane = (q << 2) | rho, rho = row group within a 16-row step, q = column block.
// The 8 lanes of each LDS.128 quarter-warp are {q0, q0+1} x {rho 0..3}:
// bank groups (FPL*q + k + rho) mod 8 are all distinct (for FPL >= 8 odd q read
// their columns rotated by 4), so every shared load is conflict free, the 4x4
// transposition happens in registers with static indexing, and each store
// instruction writes, for 4 different rows, one full 128-byte line of the row.
constexpr int kConsumerWarps = 8;
constexpr int kMaxTileRows = 256;

template <int MODE> struct ModeTraits;
// TILE rows per tile x PANEL columns per stage. The TMA unit retires roughly one
// bulk copy per ~46 cycles per SM, so 512-byte copies (128-row f32 tiles) cap the
// read rate near 3 TB/s = the 90 % of the HBM copy peak the f32 kernel reaches;
// 256-row tiles (1 KB copies), 32-column panels and 3-8 stages were all measured
// and are within 2 % (profiles/kbench_v14.jsonl), so the smallest tile stays.
// SRC = source itemsize in bytes. Modes 3/4 take 8-byte source columns (int64 /
// float64: the reference's DATA_SPEC schema, data_generation.py:56-77) - mode 3
// copies them bit for bit (plain ShufflingDataset rows), mode 4 converts every
// column to a 4-byte destination (the default torch.float feature/label types,
// torch_dataset.py:181-198) with a per-column conversion kind.
template <> struct ModeTraits<0> { static constexpr int CGROUPS = RSDL_CGROUPS_F32; static constexpr int SRC = 4; static constexpr int FPL = 4;  static constexpr int PANEL = RSDL_PANEL_F32;  static constexpr int TILE = RSDL_TILE_F32; static constexpr int STAGES = RSDL_STAGES_F32; static constexpr int MIN_CTAS = RSDL_CTAS_F32; };
template <> struct ModeTraits<1> { static constexpr int CGROUPS = RSDL_CGROUPS_F32; static constexpr int SRC = 4; static constexpr int FPL = 8;  static constexpr int PANEL = 64;  static constexpr int TILE = RSDL_TILE_F32; static constexpr int STAGES = RSDL_STAGES_F32; static constexpr int MIN_CTAS = RSDL_CTAS_F32; };
template <> struct ModeTraits<2> { static constexpr int CGROUPS = 1; static constexpr int SRC = 4; static constexpr int FPL = 16; static constexpr int PANEL = 128; static constexpr int TILE = 128; static constexpr int STAGES = 3; static constexpr int MIN_CTAS = 1; };
template <> struct ModeTraits<3> { static constexpr int CGROUPS = RSDL_CGROUPS_64; static constexpr int SRC = 8; static constexpr int FPL = 2;  static constexpr int PANEL = 32;  static constexpr int TILE = RSDL_TILE_64; static constexpr int STAGES = RSDL_STAGES_64; static constexpr int MIN_CTAS = 1; };
template <> struct ModeTraits<4> { static constexpr int CGROUPS = RSDL_CGROUPS_64; static constexpr int SRC = 8; static constexpr int FPL = 4;  static constexpr int PANEL = 32;  static constexpr int TILE = RSDL_TILE_64; static constexpr int STAGES = RSDL_STAGES_64; static constexpr int MIN_CTAS = 1; };
// Warp roles: [0, kLoaderWarps) issue the TMA loads, the next kIndexWarps
// evaluate the permutation, the rest transpose + cast + scatter.
#ifndef RSDL_INDEX_WARPS
#define RSDL_INDEX_WARPS 4
#endif
constexpr int kLoaderWarps = 4;
constexpr int kIndexWarps = RSDL_INDEX_WARPS;
template <int MODE> constexpr int kThreadsFor =
    32 * (kLoaderWarps + kIndexWarps + kConsumerWarps * ModeTraits<MODE>::CGROUPS);

// Two shared-memory tile layouts (both conflict-free for the consumers):
//  * tensor-map path: kTileRows/32 TMA boxes of [PANEL cols][32 rows] with the
//    128-byte swizzle; element (c, r) of box b lives at
//    b*PANEL*128 + c*128 + (((r%32)/4 ^ (c&7)) << 4) + (r%4)*4 bytes;
//  * 1-D bulk path (arbitrary column pointers): column c at word c*kPitchWords.
constexpr int kBoxRows = 32;

template <int MODE>
struct alignas(1024) FastSmem {
  static constexpr int PANEL = ModeTraits<MODE>::PANEL;
  static constexpr int STAGES = ModeTraits<MODE>::STAGES;
  static constexpr int TILE = ModeTraits<MODE>::TILE;
  static constexpr int kPitchWords = TILE * (ModeTraits<MODE>::SRC / 4) + 4;   // 1-D path: words per column
  static constexpr int kTileWords = ((PANEL * kPitchWords + 255) / 256) * 256;
  float tile[STAGES][kTileWords];
  unsigned long long dptr[STAGES][TILE];
  uint64_t full[STAGES];       // TMA bytes landed
  uint64_t idx_full[STAGES];   // destination pointers written
  uint64_t empty[STAGES];      // consumers done with the stage
  uint64_t turn[kIndexWarps];  // token ring: index warps publish tiles in order
};

// TAIL = the row has tail fields / a tail range (FastParams::tail_*): a separate
// instantiation, because the call to tail_step costs the kernel ~25 registers per thread
// and the tail-less kernels (the headline f32 / bf16 tables) must keep their footprint -
// what is left of the register file decides how many CTAs of the trainer's kernels can
// run next to the persistent scatter CTA.
template <int MODE, bool TAIL = false>
__global__ void __launch_bounds__(kThreadsFor<MODE>, ModeTraits<MODE>::MIN_CTAS)
scatter_tma_kernel(const __grid_constant__ FastParams p) {
  using T = ModeTraits<MODE>;
  constexpr int FPL = T::FPL;
  constexpr int PANEL = T::PANEL;
  constexpr int STAGES = T::STAGES;
  constexpr int kTileRows = T::TILE;
  constexpr int SRC = T::SRC;
  constexpr int kPitchWords = kTileRows * (SRC / 4) + 4;
  constexpr int kBoxesPerTile = kTileRows / kBoxRows;
  constexpr int PASSES = PANEL / (8 * FPL);
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128-byte TMA swizzle is a function of address bits
  const uint32_t raw = smem_u32(smem_raw);
  FastSmem<MODE>& sm = *reinterpret_cast<FastSmem<MODE>*>(smem_raw + ((1024u - (raw & 1023u)) & 1023u));
  // 1: 32-row boxes, 128B swizzle; 2: one dense box (4-byte sources only)
  const bool tmap = SRC == 4 && p.use_tmap != 0;
  const bool dense = SRC == 4 && p.use_tmap == 2;
  // Producer schedule (FastParams::sched):
  //  0  dedicated loader warps issue a stage's loads the moment it is released;
  //     every index warp owns whole tiles (kTileRows/32 rows per lane) - best for
  //     narrow tables, bf16 / 8-byte modes (index bound otherwise);
  //  1  "cooperative": the index warps share every tile (one row per lane) and
  //     issue the tile's loads themselves after publishing its index; measured
  //     10% faster for 64 x f32 (1.08 vs 1.20 ms, profiles/kbench_v16.jsonl) where
  //     the paced load issue interleaves better with the consumers' row writes.
  static_assert(kLoaderWarps == kIndexWarps, "full[] arrival count is shared by both schedules");
  const bool coop = p.sched == 1 && !tmap;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], tmap ? 1 : kLoaderWarps);   // expect_tx arrivals (== kIndexWarps)
      mbar_init(&sm.idx_full[s], coop ? kIndexWarps : 1);   // index warp(s) of the tile
      mbar_init(&sm.empty[s], kConsumerWarps);
    }
    for (int w = 0; w < kIndexWarps; ++w) mbar_init(&sm.turn[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // Pre-arm instead of waiting on "the phase before the first": every stage starts
    // released (phase 0 of empty[s] completes here), so all waiters use the plain
    // parity of their own iteration count.
    for (int s = 0; s < STAGES; ++s)
      for (int k = 0; k < kConsumerWarps; ++k) mbar_arrive(&sm.empty[s]);
  }
  __syncthreads();

  const unsigned long long num_tiles = (p.n_local + kTileRows - 1) / kTileRows;

  if (warp < kLoaderWarps) {
    if (tmap) {
    // ===== loader (tensor-map path): kBoxesPerTile TMA box loads per tile =====
    if (warp == 0 && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
      for (uint32_t panel = 0; panel < p.num_panels; ++panel) {
        mbar_wait(&sm.empty[stage], phase);
        mbar_arrive_expect_tx(&sm.full[stage], PANEL * kTileRows * 4u);   // SRC == 4 here
        if (dense) {
          // one [PANEL cols][TILE rows] box: 512-byte (or 1 KB) contiguous DRAM
          // reads per column, un-swizzled smem (2-way LDS conflicts, cheap)
          tma_load_2d(&sm.tile[stage][0], &p.tmap, static_cast<int>(tile * kTileRows),
                      static_cast<int>(panel * PANEL), &sm.full[stage]);
        } else {
#pragma unroll
          for (int b = 0; b < kBoxesPerTile; ++b)
            tma_load_2d(&sm.tile[stage][b * PANEL * kBoxRows], &p.tmap,
                        static_cast<int>(tile * kTileRows + b * kBoxRows),
                        static_cast<int>(panel * PANEL), &sm.full[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    } else if (!coop) {
    // ===== loaders (1-D bulk path): one cp.async.bulk per column of the panel =====
    // A warp issues its UBLKCPs one lane at a time (~65 cycles each) while the
    // TMA unit retires one per ~46 cycles, so every panel's columns are split
    // over the kLoaderWarps warps; all of them take part in every iteration, in
    // order, which keeps the parity test on empty[] valid.
    int stage = 0;
    uint32_t phase = 0;
    for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    for (uint32_t panel = 0; panel < p.num_panels; ++panel) {
      const uint32_t col0 = panel * PANEL;
      const uint32_t ncols = min(static_cast<uint32_t>(PANEL), p.num_cols - col0);
      const uint32_t per = (ncols + kLoaderWarps - 1) / kLoaderWarps;
      const uint32_t c_lo = min(warp * per, ncols), c_hi = min(c_lo + per, ncols);
      mbar_wait(&sm.empty[stage], phase);
      if (lane == 0) mbar_arrive_expect_tx(&sm.full[stage], (c_hi - c_lo) * kTileRows * SRC);
      __syncwarp();
      for (uint32_t c = c_lo + lane; c < c_hi; c += 32) {
        const uint8_t* src = p.cols[col0 + c] + tile * static_cast<unsigned long long>(kTileRows * SRC);
        // 8-byte sources: odd column blocks sit 16 bytes later (fits exactly in
        // the 16-byte column padding) so that a quarter-warp's two column
        // blocks fall in different 16-byte bank groups - see the consumer.
        const uint32_t shift = (SRC == 8) ? (((c / FPL) & 1u) << 2) : 0u;
        tma_load_1d(&sm.tile[stage][c * kPitchWords + shift], src, kTileRows * SRC, &sm.full[stage]);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    }
  } else if (warp < kLoaderWarps + kIndexWarps) {
    // ===== index warps: the tile's shared permutation index =====
    // Each index warp owns whole tiles (round-robin), so kIndexWarps tiles are
    // indexed concurrently, and a lane evaluates the Feistel network for
    // kTileRows/32 rows at once (independent dependency chains). With all index
    // warps cooperating on one tile (one row per lane) the serial chain
    // "cycle-walk (max over the lanes) -> publish -> issue loads" bounded the
    // tile rate at ~1.2 us regardless of the row width, i.e. narrow tables ran at
    // a third of the HBM rate (profiles/README.md, v8 -> v9).
    //
    // Publishing stays strictly in tile order through a token ring of mbarriers
    // (turn[w] is completed by warp w-1 when it is done with its tile): the
    // parity test on empty[] is only meaningful for a waiter that is at most one
    // phase ahead of the barrier, which free-running warps would not guarantee.
    // The expensive part - the Feistel evaluations - happens before the token
    // is awaited.
    const int w = warp - kLoaderWarps;
    if (coop) {
      // schedule 1: all index warps cooperate on every tile and load it
      constexpr int kRows = kTileRows / (32 * kIndexWarps);
      const int r = w * 32 + lane;
      int stage = 0;
      uint32_t phase = 0;
      for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        unsigned long long prev[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
          const unsigned long long lr = tile * kTileRows + r + i * (32 * kIndexWarps);
          prev[i] = (lr < p.n_local)
              ? dest_pointer(p.global_offset + lr, p.key, p.plan, p.dst, p.row_pitch) : 0ull;
        }
        for (uint32_t panel = 0; panel < p.num_panels; ++panel) {
          mbar_wait(&sm.empty[stage], phase);
#pragma unroll
          for (int i = 0; i < kRows; ++i) sts64(&sm.dptr[stage][r + i * (32 * kIndexWarps)], prev[i]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.idx_full[stage]);
          rc_arrive(stage, 32 * (kIndexWarps + kConsumerWarps));
          {
            const uint32_t col0 = panel * PANEL;
            const uint32_t ncols = min(static_cast<uint32_t>(PANEL), p.num_cols - col0);
            const uint32_t per = (ncols + kIndexWarps - 1) / kIndexWarps;
            const uint32_t c_lo = min(w * per, ncols), c_hi = min(c_lo + per, ncols);
            if (lane == 0) mbar_arrive_expect_tx(&sm.full[stage], (c_hi - c_lo) * kTileRows * SRC);
            __syncwarp();
            for (uint32_t c = c_lo + lane; c < c_hi; c += 32) {
              const uint8_t* src = p.cols[col0 + c] + tile * static_cast<unsigned long long>(kTileRows * SRC);
              const uint32_t shift = (SRC == 8) ? (((c / FPL) & 1u) << 2) : 0u;
              tma_load_1d(&sm.tile[stage][c * kPitchWords + shift], src, kTileRows * SRC, &sm.full[stage]);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else {
    constexpr int RPL = kTileRows / 32;        // rows per lane
    int stage = 0;
    uint32_t phase = 0;
    int owner = 0;
    uint32_t round = 0;                        // tiles this warp has published
    for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const bool mine = owner == w;
      if (++owner == kIndexWarps) owner = 0;
      if (!mine) {
        for (uint32_t panel = 0; panel < p.num_panels; ++panel)
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        continue;
      }
      // evaluated once per (row, tile); every column panel of the tile reuses it
      unsigned long long prev[RPL];
      {
        unsigned long long x[RPL];
        bool valid[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          const unsigned long long lr = tile * kTileRows + lane + 32 * i;
          valid[i] = lr < p.n_local;
          x[i] = valid[i] ? p.global_offset + lr : 0ull;
        }
        if (p.key.n > 1) {
#pragma unroll
          for (int i = 0; i < RPL; ++i) x[i] = rsdl_feistel(x[i], p.key);
          // Cycle-walk the rows that left [0, n) (same result as rsdl_permute).
          // About a quarter of the rows need it and very few need it twice, so
          // after the first (RPL-way interleaved) evaluation each round walks
          // ONE row per lane - the lane's first row still outside - instead of
          // re-evaluating all RPL: the kernel is issue bound for narrow tables
          // (profiles/README.md, v11-v13), and the round count (max over the
          // warp) stays about the same while a round costs 1/RPL of the work.
          while (true) {
            int sel = -1;
#pragma unroll
            for (int i = RPL - 1; i >= 0; --i) sel = (x[i] >= p.key.n) ? i : sel;
            if (!__any_sync(0xffffffffu, sel >= 0)) break;
            unsigned long long xs = x[0];
#pragma unroll
            for (int i = 1; i < RPL; ++i) xs = (sel == i) ? x[i] : xs;
            const unsigned long long y = rsdl_feistel(xs, p.key);
#pragma unroll
            for (int i = 0; i < RPL; ++i) x[i] = (sel == i) ? y : x[i];
          }
        }
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          uint32_t trainer;
          unsigned long long slot;
          rsdl_position_to_dest(x[i], p.plan, &trainer, &slot);
          prev[i] = (valid[i] && slot >= p.plan.slot_lo && slot < p.plan.slot_hi)
              ? reinterpret_cast<unsigned long long>(p.dst[trainer]) + slot * p.row_pitch
              : 0ull;
        }
      }
      // the token: warp 0 holds it initially, so its first round does not wait at all
      // (round r of warp 0 waits for completion r-1 of turn[0], warp w > 0 for
      // completion r of turn[w])
      if (w != 0) mbar_wait(&sm.turn[w], round & 1u);
      else if (round != 0) mbar_wait(&sm.turn[0], (round - 1u) & 1u);
      ++round;
      for (uint32_t panel = 0; panel < p.num_panels; ++panel) {
        mbar_wait(&sm.empty[stage], phase);
#pragma unroll
        for (int i = 0; i < RPL; ++i) sts64(&sm.dptr[stage][lane + 32 * i], prev[i]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.idx_full[stage]);
        rc_arrive(stage, 32 * (1 + kConsumerWarps));
        if (lane == 0 && panel + 1 == p.num_panels) mbar_arrive(&sm.turn[(w + 1) % kIndexWarps]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    }  // schedule 0
  } else {
    // ===== consumers: transpose + cast + scatter =====
    const int cwarp = (warp - kLoaderWarps - kIndexWarps) % kConsumerWarps;
    const int cgroup = (warp - kLoaderWarps - kIndexWarps) / kConsumerWarps;
    const int rho = lane & 3;
    const int q = lane >> 2;
    int stage = 0;
    uint32_t phase = 0;
    int turn = 0;                              // which consumer group owns this iteration
    for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    for (uint32_t panel = 0; panel < p.num_panels; ++panel) {
      if (T::CGROUPS > 1) {
        const bool mine = turn == cgroup;
        if (++turn == T::CGROUPS) turn = 0;
        if (!mine) {
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
      }
      const uint32_t col0 = panel * PANEL;
      const uint32_t ncols = min(static_cast<uint32_t>(PANEL), p.num_cols - col0);
      mbar_wait(&sm.idx_full[stage], phase);
      rc_sync(stage, 32 * ((coop ? kIndexWarps : 1) + kConsumerWarps));
      mbar_wait(&sm.full[stage], phase);
      const float* A = sm.tile[stage];
      for (int step = cwarp; step < kTileRows / 16; step += kConsumerWarps) {
        const int rg = step * 4 + rho;            // 4-row group inside the tile
        unsigned long long d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = lds64(&sm.dptr[stage][rg * 4 + j]);
        if constexpr (SRC == 8) {
          // 8-byte sources. Column c lives at word c*kPitchWords (+4 for odd column
          // blocks); rows 4rg..4rg+3 of it are two 16-byte groups whose bank
          // group is (c + [block odd] + 2rg + h) mod 8 - distinct for the 8 lanes
          // {q0, q0+1} x {rho 0..3} of a quarter-warp.
#pragma unroll
          for (int pass = 0; pass < PASSES; ++pass) {
            const int cb = q + 8 * pass;
            const uint32_t f0 = cb * FPL;
            // columns past the last one are written as zeros up to write_end (the
            // row's padding), so that no 32-byte sector is left partially written
            if (f0 >= ncols && (col0 + f0) * (16u / FPL) >= p.write_end) continue;
            unsigned long long v[FPL][4];
#pragma unroll
            for (int k = 0; k < FPL; ++k) {
              const uint32_t f = f0 + k;
              if (f < ncols) {
                const float* w = A + f * kPitchWords + ((cb & 1) << 2) + rg * 8;
                const float4 lo = lds128(w), hi = lds128(w + 4);
                v[k][0] = (static_cast<unsigned long long>(__float_as_uint(lo.y)) << 32) | __float_as_uint(lo.x);
                v[k][1] = (static_cast<unsigned long long>(__float_as_uint(lo.w)) << 32) | __float_as_uint(lo.z);
                v[k][2] = (static_cast<unsigned long long>(__float_as_uint(hi.y)) << 32) | __float_as_uint(hi.x);
                v[k][3] = (static_cast<unsigned long long>(__float_as_uint(hi.w)) << 32) | __float_as_uint(hi.z);
              } else {
                v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0ull;
              }
            }
            if (MODE == 3) {
              const uint32_t off = (col0 + f0) * 8u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (d[j])
                  stg128(reinterpret_cast<void*>(d[j] + off), static_cast<uint32_t>(v[0][j]),
                         static_cast<uint32_t>(v[0][j] >> 32), static_cast<uint32_t>(v[1][j]),
                         static_cast<uint32_t>(v[1][j] >> 32));
              }
            } else {
              // kind 0: int64 -> f32 (RNE), 1: float64 -> f32 (RNE), 2: int64 -> int32.
              // The table is padded to a multiple of 4 entries: one aligned word
              // holds this lane's 4 kinds. All-int64->f32 (the common case) runs
              // branch free; mixed blocks compute both conversions and select.
              const uint32_t k4 = (f0 < ncols)
                  ? __ldg(reinterpret_cast<const uint32_t*>(p.kinds + col0 + f0)) : 0u;
              uint32_t o[FPL][4];
              if (k4 == 0u) {
#pragma unroll
                for (int k = 0; k < FPL; ++k)
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    o[k][j] = __float_as_uint(__ll2float_rn(static_cast<long long>(v[k][j])));
              } else {
#pragma unroll
                for (int k = 0; k < FPL; ++k) {
                  const uint32_t kind = (k4 >> (8 * k)) & 0xFFu;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t fi = __float_as_uint(__ll2float_rn(static_cast<long long>(v[k][j])));
                    const uint32_t fd = __float_as_uint(
                        __double2float_rn(__longlong_as_double(static_cast<long long>(v[k][j]))));
                    o[k][j] = kind == 0u ? fi : (kind == 1u ? fd : static_cast<uint32_t>(v[k][j]));
                  }
                }
              }
              const uint32_t off = (col0 + f0) * 4u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (d[j]) stg128(reinterpret_cast<void*>(d[j] + off), o[0][j], o[1][j], o[2 % FPL][j], o[3 % FPL][j]);
              }
            }
          }
        } else {
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
          const int cb = q + 8 * pass;            // column block inside the panel
          const uint32_t f0 = cb * FPL;
          // (columns past the last one: zeros up to write_end, see the 8-byte path)
          const bool active = f0 < ncols || (MODE != 2 && (col0 + f0) * (16u / FPL) < p.write_end);
          // (fp8 mode exchanges amax with the partner lane: nobody may skip)
          if (MODE != 2 && !active) continue;
          float v[FPL][4];
#pragma unroll
          for (int k = 0; k < FPL; ++k) {
            // odd column blocks read rotated by 4 so that a quarter-warp's two
            // column blocks never share a bank group when FPL is a multiple of 8
            const int kk = (FPL >= 8) ? ((k + 4 * (q & 1)) & (FPL - 1)) : k;
            const uint32_t f = f0 + kk;
            const float* src_word = dense
                ? A + f * kTileRows + rg * 4
                : (tmap ? A + (rg >> 3) * (PANEL * kBoxRows) + f * kBoxRows +
                              ((((rg & 7) ^ (f & 7))) << 2)
                        : A + f * kPitchWords + rg * 4);
            float4 x = (f < ncols) ? lds128(src_word) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (FPL >= 8 && (q & 1)) {
              // undo the rotation with static register indices
              v[(k + 4) & (FPL - 1)][0] = x.x; v[(k + 4) & (FPL - 1)][1] = x.y;
              v[(k + 4) & (FPL - 1)][2] = x.z; v[(k + 4) & (FPL - 1)][3] = x.w;
            } else {
              v[k][0] = x.x; v[k][1] = x.y; v[k][2] = x.z; v[k][3] = x.w;
            }
          }
          if (MODE == 0) {
            const uint32_t off = (col0 + f0) * 4u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (d[j])
                stg128(reinterpret_cast<void*>(d[j] + off), __float_as_uint(v[0][j]),
                       __float_as_uint(v[1][j]), __float_as_uint(v[2][j]),
                       __float_as_uint(v[3][j]));
            }
          } else if (MODE == 1) {
            const uint32_t off = (col0 + f0) * 2u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (d[j])
                stg128(reinterpret_cast<void*>(d[j] + off), pack_bf16x2(v[0][j], v[1][j]),
                       pack_bf16x2(v[2][j], v[3][j]), pack_bf16x2(v[4][j], v[5][j]),
                       pack_bf16x2(v[6][j], v[7][j]));
            }
          } else {
            // Block-scaled e4m3: 32-element blocks = this lane + its q^1 partner.
            const uint32_t off = col0 + f0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float amax = 0.f;
#pragma unroll
              for (int k = 0; k < FPL; ++k) {
                float a = fabsf(v[k][j]);
                amax = (a > amax) ? a : amax;      // NaN never wins: matches golden
              }
              amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
              const uint32_t bits = __float_as_uint(amax * (1.0f / 448.0f));
              int e = static_cast<int>((bits >> 23) & 0xFF) - 127 + ((bits & 0x7FFFFF) ? 1 : 0);
              e = max(-126, min(127, e));
              const float inv = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);
              uint32_t w[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                w[g] = f32x2_to_e4m3x2(v[4 * g + 0][j] * inv, v[4 * g + 1][j] * inv) |
                       (f32x2_to_e4m3x2(v[4 * g + 2][j] * inv, v[4 * g + 3][j] * inv) << 16);
              }
              // UE8M0 scale bytes of the panel's four 32-column blocks (held by the
              // lanes q = 0, 2, 4, 6 of this row group) are gathered into ONE aligned
              // 4-byte store by the q = 0 lane: single-byte stores to a peer are one
              // NVLink write transaction each and made the fp8 epilogue slower over
              // the link than f32 (2.89 vs 2.43 ms at N = 2). Blocks past the last
              // column contribute 0, which is what the zero-initialised padding holds.
              const uint32_t sb = (active && amax > 0.f) ? static_cast<uint32_t>(e + 127) : 0u;
              const uint32_t s1 = __shfl_sync(0xffffffffu, sb, rho + 8);
              const uint32_t s2 = __shfl_sync(0xffffffffu, sb, rho + 16);
              const uint32_t s3 = __shfl_sync(0xffffffffu, sb, rho + 24);
              if (d[j] && active) {
                stg128(reinterpret_cast<void*>(d[j] + off), w[0], w[1], w[2], w[3]);
                if (q == 0)
                  *reinterpret_cast<uint32_t*>(d[j] + p.scale_offset + (off >> 5)) =
                      sb | (s1 << 8) | (s2 << 16) | (s3 << 24);
              }
            }
          }
        }
        }  // SRC == 4
        // ---- tail step: bytes [tail_lo, tail_hi) of every row (see tail_step) - once per
        // tile, with the last column panel; a real call, so that its registers (type
        // dispatch, fp64 conversions) do not count against the main loop's
        if constexpr (TAIL) {
          if (panel + 1 == p.num_panels)
            tail_step(&p, d[0], d[1], d[2], d[3],
                      tile * static_cast<unsigned long long>(kTileRows) + rg * 4, q);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  // Make this CTA's peer stores visible system-wide before the grid retires
  // (the completion flag is published by a later kernel on the same stream).
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// K2 generic path: arbitrary per-field dtype / width / cast
// ---------------------------------------------------------------------------
__device__ __forceinline__ void store_cast(uint8_t* dst, uint32_t dst_code, bool is_int, long long iv,
                                           double fv, float sv, bool from_f64) {
  // `sv` is the exact fp32 value when the source is <= 32-bit float; `fv` the
  // fp64 value when the source is float64; `iv` the integer value otherwise.
  switch (dst_code) {
    case DT_F32: {
      float o = is_int ? static_cast<float>(iv) : (from_f64 ? __double2float_rn(fv) : sv);
      *reinterpret_cast<float*>(dst) = o;
      break;
    }
    case DT_F64: {
      double o = is_int ? static_cast<double>(iv) : (from_f64 ? fv : static_cast<double>(sv));
      // 8-byte fields are 8-aligned in the row but the staging pitch is odd in
      // words: store as two halves.
      unsigned long long b = __double_as_longlong(o);
      reinterpret_cast<uint32_t*>(dst)[0] = static_cast<uint32_t>(b);
      reinterpret_cast<uint32_t*>(dst)[1] = static_cast<uint32_t>(b >> 32);
      break;
    }
    case DT_F16: {
      __half o = is_int ? __ll2half_rn(iv) : (from_f64 ? __double2half(fv) : __float2half_rn(sv));
      *reinterpret_cast<__half*>(dst) = o;
      break;
    }
    case DT_BF16: {
      float f = is_int ? static_cast<float>(iv) : (from_f64 ? __double2float_rn(fv) : sv);
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(f);
      break;
    }
    case DT_FP8: {
      float f = is_int ? static_cast<float>(iv) : (from_f64 ? __double2float_rn(fv) : sv);
      *dst = static_cast<uint8_t>(f32_to_e4m3(f));
      break;
    }
    case DT_BOOL: {
      bool nz = is_int ? (iv != 0) : (from_f64 ? (fv != 0.0) : (sv != 0.f));
      *dst = nz ? 1 : 0;
      break;
    }
    default: {
      long long o = is_int ? iv : (from_f64 ? static_cast<long long>(fv) : static_cast<long long>(sv));
      switch (dst_code) {
        case DT_U8: case DT_I8: *dst = static_cast<uint8_t>(o); break;
        case DT_I16: *reinterpret_cast<int16_t*>(dst) = static_cast<int16_t>(o); break;
        case DT_I32: *reinterpret_cast<int32_t*>(dst) = static_cast<int32_t>(o); break;
        default:
          reinterpret_cast<uint32_t*>(dst)[0] = static_cast<uint32_t>(o);
          reinterpret_cast<uint32_t*>(dst)[1] = static_cast<uint32_t>(static_cast<unsigned long long>(o) >> 32);
      }
    }
  }
}

__device__ __forceinline__ void load_src(const uint8_t* src, uint32_t code, bool* is_int,
                                         long long* iv, double* fv, float* sv, bool* from_f64) {
  *is_int = false; *from_f64 = false; *iv = 0; *fv = 0.0; *sv = 0.f;
  switch (code) {
    case DT_U8: case DT_BOOL: *is_int = true; *iv = *src; break;
    case DT_I8: *is_int = true; *iv = *reinterpret_cast<const int8_t*>(src); break;
    case DT_I16: *is_int = true; *iv = *reinterpret_cast<const int16_t*>(src); break;
    case DT_I32: *is_int = true; *iv = *reinterpret_cast<const int32_t*>(src); break;
    case DT_I64: *is_int = true; *iv = *reinterpret_cast<const long long*>(src); break;
    case DT_F16: *sv = __half2float(*reinterpret_cast<const __half*>(src)); break;
    case DT_BF16: *sv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(src)); break;
    case DT_F32: *sv = *reinterpret_cast<const float*>(src); break;
    case DT_F64: *from_f64 = true; *fv = *reinterpret_cast<const double*>(src); break;
    default: break;
  }
}

__global__ void __launch_bounds__(256) scatter_generic_kernel(const GenericParams p) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const uint32_t R = p.rows_per_block;
  // Only the byte range this launch owns is staged (a 8-byte label next to a
  // 24 KB image row must not cost 24 KB of shared memory per row).
  const uint32_t span = p.write_hi - p.write_lo;
  const uint32_t spitch = span + 4;            // odd word count: conflict-free columns
  unsigned long long* dptr = reinterpret_cast<unsigned long long*>(gsm);
  uint8_t* stage = gsm + R * sizeof(unsigned long long);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;

  for (unsigned long long row0 = static_cast<unsigned long long>(blockIdx.x) * R; row0 < p.n_local;
       row0 += static_cast<unsigned long long>(gridDim.x) * R) {
    const uint32_t rows = static_cast<uint32_t>(min(static_cast<unsigned long long>(R), p.n_local - row0));
    for (uint32_t r = threadIdx.x; r < rows; r += blockDim.x)
      dptr[r] = dest_pointer(p.global_offset + row0 + r, p.key, p.plan, p.dst, p.row_pitch);
    // zero the staging rows so padding bytes are deterministic
    for (uint32_t w = threadIdx.x; w < rows * (spitch / 4); w += blockDim.x)
      reinterpret_cast<uint32_t*>(stage)[w] = 0;
    __syncthreads();
    // phase 1: (field, row) work items; lanes run along rows => coalesced loads
    const uint32_t work = p.num_fields * rows;
    for (uint32_t idx = threadIdx.x; idx < work; idx += blockDim.x) {
      const uint32_t fi = idx / rows;
      const uint32_t r = idx - fi * rows;
      const FieldDev f = p.fields[fi];
      const uint32_t ss = rsdl_itemsize(f.src_code), ds = rsdl_itemsize(f.dst_code);
      const uint8_t* src = f.src + (row0 + r) * static_cast<unsigned long long>(ss) * f.width;
      uint8_t* out = stage + r * spitch + (f.dst_off - p.write_lo);
      for (uint32_t w = 0; w < f.width; ++w) {
        bool is_int, from_f64; long long iv; double fv; float sv;
        load_src(src + w * ss, f.src_code, &is_int, &iv, &fv, &sv, &from_f64);
        store_cast(out + w * ds, f.dst_code, is_int, iv, fv, sv, from_f64);
      }
    }
    __syncthreads();
    // phase 2: one warp per row, 128 B contiguous per store instruction
    const uint32_t nwords = span / 4;
    for (uint32_t r = warp; r < rows; r += nwarps) {
      if (dptr[r] == 0ull) continue;            // row belongs to another pass
      uint32_t* drow = reinterpret_cast<uint32_t*>(dptr[r] + p.write_lo);
      const uint32_t* srow = reinterpret_cast<const uint32_t*>(stage + r * spitch);
      for (uint32_t w = lane; w < nwords; w += 32) drow[w] = srow[w];
    }
    __syncthreads();
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// K2 wide path: list-valued columns ([N, width] row-major, e.g. images)
// ---------------------------------------------------------------------------
// The source is already row-major, so the scatter is a permuted row copy with an
// optional cast: one warp per row, 16-byte vector loads and stores on both sides.
__global__ void __launch_bounds__(256) scatter_wide_kernel(const WideParams p) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp0 =
      (blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x) >> 5;
  const unsigned long long nwarps = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 5;
  const uint32_t ss = rsdl_itemsize(p.src_code), ds = rsdl_itemsize(p.dst_code);
  const unsigned long long src_row_bytes = static_cast<unsigned long long>(p.width) * ss;
  const bool vec_copy = (p.src_code == p.dst_code) && (src_row_bytes % 16 == 0) &&
                        (p.dst_off % 16 == 0) && (p.row_pitch % 16 == 0) &&
                        ((reinterpret_cast<unsigned long long>(p.src) & 15) == 0);
  const bool vec_bf16 = (p.src_code == DT_F32 && p.dst_code == DT_BF16) && (p.width % 8 == 0) &&
                        (p.dst_off % 16 == 0) && (p.row_pitch % 16 == 0) &&
                        ((reinterpret_cast<unsigned long long>(p.src) & 15) == 0);
  // uint8 pixels -> f32 / f16 / bf16 (image columns stored as bytes, delivered in the
  // training dtype): 16 pixels per lane per step, one 16-byte load, 2-4 16-byte stores
  const bool vec_u8 = (p.src_code == DT_U8) &&
                      (p.dst_code == DT_F32 || p.dst_code == DT_F16 || p.dst_code == DT_BF16) &&
                      (p.width % 16 == 0) && (p.dst_off % 16 == 0) && (p.row_pitch % 16 == 0) &&
                      ((reinterpret_cast<unsigned long long>(p.src) & 15) == 0);
  for (unsigned long long r = warp0; r < p.n_local; r += nwarps) {
    const unsigned long long d0 =
        dest_pointer(p.global_offset + r, p.key, p.plan, p.dst, p.row_pitch);
    if (d0 == 0ull) continue;                   // row belongs to another pass
    const unsigned long long d = d0 + p.dst_off;
    const uint8_t* src = p.src + r * src_row_bytes;
    if (vec_copy) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      const uint32_t nvec = static_cast<uint32_t>(src_row_bytes / 16);
      for (uint32_t v = lane; v < nvec; v += 32) d4[v] = __ldg(s4 + v);
    } else if (vec_bf16) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      const uint32_t nvec = p.width / 8;
      for (uint32_t v = lane; v < nvec; v += 32) {
        const float4 a = __ldg(s4 + 2 * v), b = __ldg(s4 + 2 * v + 1);
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
        d4[v] = o;
      }
    } else if (vec_u8) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      const uint32_t nvec = p.width / 16;
      for (uint32_t v = lane; v < nvec; v += 32) {
        const uint4 raw = __ldg(s4 + v);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        if (p.dst_code == DT_F32) {
          uint4* d4 = reinterpret_cast<uint4*>(d) + 4ull * v;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = __float_as_uint(static_cast<float>(w[g] & 0xFFu));
            o.y = __float_as_uint(static_cast<float>((w[g] >> 8) & 0xFFu));
            o.z = __float_as_uint(static_cast<float>((w[g] >> 16) & 0xFFu));
            o.w = __float_as_uint(static_cast<float>(w[g] >> 24));
            d4[g] = o;
          }
        } else {
          // 0..255 is exact in both half formats
          uint32_t h[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float a = static_cast<float>(w[g] & 0xFFu), b = static_cast<float>((w[g] >> 8) & 0xFFu);
            const float c = static_cast<float>((w[g] >> 16) & 0xFFu), e = static_cast<float>(w[g] >> 24);
            if (p.dst_code == DT_BF16) {
              h[2 * g] = pack_bf16x2(a, b);
              h[2 * g + 1] = pack_bf16x2(c, e);
            } else {
              const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, e);
              h[2 * g] = *reinterpret_cast<const uint32_t*>(&lo);
              h[2 * g + 1] = *reinterpret_cast<const uint32_t*>(&hi);
            }
          }
          uint4* d4 = reinterpret_cast<uint4*>(d) + 2ull * v;
          d4[0] = make_uint4(h[0], h[1], h[2], h[3]);
          d4[1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
      }
    } else {
      for (uint32_t e = lane; e < p.width; e += 32) {
        bool is_int, from_f64; long long iv; double fv; float sv;
        load_src(src + static_cast<unsigned long long>(e) * ss, p.src_code, &is_int, &iv, &fv, &sv, &from_f64);
        uint8_t* out = reinterpret_cast<uint8_t*>(d) + static_cast<unsigned long long>(e) * ds;
        if (ds == 8) {
          // store_cast writes 8-byte values as two words (alignment-agnostic)
          store_cast(out, p.dst_code, is_int, iv, fv, sv, from_f64);
        } else {
          store_cast(out, p.dst_code, is_int, iv, fv, sv, from_f64);
        }
      }
    }
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// K1 standalone: positions of local rows (tests, NCCL baseline)
// ---------------------------------------------------------------------------
__global__ void perm_positions_kernel(PermKeyDev key, PlanDev plan, unsigned long long global_offset,
                                      unsigned long long n_local, int32_t* trainer, long long* slot) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
       i < n_local; i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    unsigned long long pos = rsdl_permute(global_offset + i, key);
    uint32_t t; unsigned long long s;
    rsdl_position_to_dest(pos, plan, &t, &s);
    trainer[i] = static_cast<int32_t>(t);
    slot[i] = static_cast<long long>(s);
  }
}

// ---------------------------------------------------------------------------
// K5 local row placement (NCCL baseline's gather and final placement stages):
//   dst_base + dst_off[i]  <-  rows[(src_idx ? src_idx[i] : i)]     (dst_off < 0: skip)
// ---------------------------------------------------------------------------
__global__ void place_rows_kernel(const uint8_t* rows, const long long* src_idx,
                                  const long long* dst_off, unsigned long long n,
                                  uint32_t pitch, uint8_t* dst) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x) >> 5;
  const unsigned long long nwarps = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 5;
  const uint32_t vecs = pitch / 16;
  for (unsigned long long i = warp; i < n; i += nwarps) {
    const long long off = dst_off[i];
    if (off < 0) continue;                      // padding entry of a fixed-size exchange block
    const unsigned long long si = src_idx ? static_cast<unsigned long long>(src_idx[i]) : i;
    const uint4* s = reinterpret_cast<const uint4*>(rows + si * pitch);
    uint4* d = reinterpret_cast<uint4*>(dst + off);
    for (uint32_t v = lane; v < vecs; v += 32) d[v] = s[v];
  }
}

// ---------------------------------------------------------------------------
// K13 key_checksum + batch reduction (consumer-side verification / bench sink)
// ---------------------------------------------------------------------------
__global__ void key_checksum_kernel(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                                    uint32_t key_off, unsigned long long* out /*[2]: sum, xor*/) {
  unsigned long long s = 0, x = 0;
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
       i < rows; i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    unsigned long long k = *reinterpret_cast<const unsigned long long*>(packed + i * pitch + key_off);
    s += k;
    // xor of a mixed key: order independent, sensitive to duplicates
    unsigned long long z = k + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    x ^= z ^ (z >> 31);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    x ^= __shfl_xor_sync(0xffffffffu, x, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], s);
    atomicXor(&out[1], x);
  }
}

__global__ void batch_sum_f32_kernel(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                                     uint32_t off, double* out) {
  double s = 0.0;
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
       i < rows; i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    s += static_cast<double>(*reinterpret_cast<const float*>(packed + i * pitch + off));
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// Full-batch sink: fp64 sum of every fp32 word of a packed batch (reads the
// whole batch once, 16 B per lane - the "trainer touched every byte" proof).
__global__ void __launch_bounds__(256) batch_sum_all_f32_kernel(const float4* data, unsigned long long nvec, double* out) {
  // 4 independent 16-byte loads in flight per thread; the launcher keeps the grid at
  // 4 CTAs per SM (1024 of the SM's 2048 threads) so that a co-running persistent
  // kernel on another stream - the next epoch's scatter - can still be scheduled.
  double s = 0.0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    const float4 a = __ldg(data + i), b = __ldg(data + i + stride);
    const float4 c = __ldg(data + i + 2 * stride), d = __ldg(data + i + 3 * stride);
    s += (static_cast<double>(a.x) + static_cast<double>(a.y)) + (static_cast<double>(a.z) + static_cast<double>(a.w));
    s += (static_cast<double>(b.x) + static_cast<double>(b.y)) + (static_cast<double>(b.z) + static_cast<double>(b.w));
    s += (static_cast<double>(c.x) + static_cast<double>(c.y)) + (static_cast<double>(c.z) + static_cast<double>(c.w));
    s += (static_cast<double>(d.x) + static_cast<double>(d.y)) + (static_cast<double>(d.z) + static_cast<double>(d.w));
  }
  for (; i < nvec; i += stride) {
    const float4 v = __ldg(data + i);
    s += (static_cast<double>(v.x) + static_cast<double>(v.y)) + (static_cast<double>(v.z) + static_cast<double>(v.w));
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(out, t);
  }
}

// ---------------------------------------------------------------------------
// Probe: TMA bulk *stores* (cp.async.bulk shared::cta -> global, SASS UBLKCP with a
// shared source) of packed rows to random slots of a local or peer buffer.
// VERDICT r1 asked for this variant of the scatter epilogue to be measured: one
// bulk store per packed row from a row-major smem tile instead of per-lane
// STG.E.128. The probe isolates exactly that store stream (the smem tile is filled
// once; every CTA then streams `rows_per_cta` row stores, one per lane per round,
// with at most 2 bulk groups in flight per lane) so its rate can be compared with
// the scatter kernel's achieved store rate at the same row size. See
// profiles/README.md "bulk stores" for the numbers and why the epilogue keeps STG.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bulk_store_probe_kernel(uint8_t* dst, unsigned long long slots,
                                                               uint32_t row_bytes,
                                                               unsigned long long rows_per_cta) {
  extern __shared__ __align__(128) uint8_t psm[];
  const uint32_t tile_rows = 32768u / row_bytes;          // 32 KB staging tile
  for (uint32_t w = threadIdx.x; w < 32768u / 4; w += blockDim.x)
    reinterpret_cast<uint32_t*>(psm)[w] = w * 2654435761u + blockIdx.x;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * rows_per_cta;
  for (unsigned long long r = threadIdx.x; r < rows_per_cta; r += blockDim.x) {
    // pseudo-random destination slot (a bijection is not needed for a rate probe)
    unsigned long long z = (base + r) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 29;
    const unsigned long long slot = z % slots;
    const uint32_t src = smem_u32(psm + static_cast<size_t>(r % tile_rows) * row_bytes);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     dst + slot * row_bytes),
                 "r"(src), "r"(row_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence_system();
}

void launch_bulk_store_probe(uint8_t* dst, unsigned long long slots, uint32_t row_bytes,
                             unsigned long long total_rows, int grid, cudaStream_t stream) {
  if (row_bytes % 16 || row_bytes == 0 || row_bytes > 32768u)
    throw std::runtime_error("bulk_store_probe: row_bytes must be a multiple of 16, <= 32 KB");
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(bulk_store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    configured = true;
  }
  bulk_store_probe_kernel<<<grid, 256, 32768, stream>>>(dst, slots, row_bytes,
                                                        (total_rows + grid - 1) / grid);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("bulk_store_probe: ") + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------
// K11 device ring signalling: epoch-tagged flags in (peer) HBM
// ---------------------------------------------------------------------------
__global__ void signal_flags_kernel(FlagTargets t, uint32_t value) {
  if (threadIdx.x < t.count) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.ptr[threadIdx.x]), "r"(value) : "memory");
  }
}

// Stream-side wait: spins (with back-off) until every flag >= value, or the
// timeout expires, in which case *error is set instead of hanging the stream.
__global__ void wait_flags_kernel(const uint32_t* flags, uint32_t count, uint32_t value,
                                  unsigned long long timeout_ns, uint32_t* error) {
  if (threadIdx.x >= count) return;
  unsigned long long start;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
  uint32_t ns = 64;
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
    if (static_cast<int32_t>(v - value) >= 0) break;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (timeout_ns && now - start > timeout_ns) {
      atomicExch(error, 1u + threadIdx.x);
      break;
    }
    __nanosleep(ns);
    if (ns < 4096) ns <<= 1;
  }
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
template <int MODE, bool TAIL>
static void launch_fast_variant(const FastParams& p, int grid, cudaStream_t stream) {
  const size_t smem = sizeof(FastSmem<MODE>) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(scatter_tma_kernel<MODE, TAIL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured = true;
  }
  scatter_tma_kernel<MODE, TAIL><<<grid, kThreadsFor<MODE>, smem, stream>>>(p);
}

template <int MODE>
static void launch_fast_mode(const FastParams& p, int grid, cudaStream_t stream) {
  if (p.tail_hi > p.tail_lo) launch_fast_variant<MODE, true>(p, grid, stream);
  else launch_fast_variant<MODE, false>(p, grid, stream);
}

#define RSDL_MODE_SWITCH(mode, EXPR)                     \
  switch (mode) {                                        \
    case 0: { using M = ModeTraits<0>; return EXPR; }    \
    case 1: { using M = ModeTraits<1>; return EXPR; }    \
    case 2: { using M = ModeTraits<2>; return EXPR; }    \
    case 3: { using M = ModeTraits<3>; return EXPR; }    \
    case 4: { using M = ModeTraits<4>; return EXPR; }    \
    default: throw std::runtime_error("bad fast scatter mode"); \
  }
int fast_panel_cols(int mode) { RSDL_MODE_SWITCH(mode, M::PANEL) }
int fast_ctas_per_sm(int mode) { RSDL_MODE_SWITCH(mode, M::MIN_CTAS) }
int fast_tile_rows(int mode) { RSDL_MODE_SWITCH(mode, M::TILE) }
int fast_src_itemsize(int mode) { RSDL_MODE_SWITCH(mode, M::SRC) }
int fast_max_tile_rows() { return kMaxTileRows; }

void launch_scatter_fast(const FastParams& p, int mode, int grid, cudaStream_t stream) {
  if (p.n_local == 0) return;
  switch (mode) {
    case 0: launch_fast_mode<0>(p, grid, stream); break;
    case 1: launch_fast_mode<1>(p, grid, stream); break;
    case 2: launch_fast_mode<2>(p, grid, stream); break;
    case 3: launch_fast_mode<3>(p, grid, stream); break;
    case 4:
      if (p.kinds == nullptr) throw std::runtime_error("scatter mode 4 needs the per-column kinds table");
      launch_fast_mode<4>(p, grid, stream);
      break;
    default: throw std::runtime_error("bad fast scatter mode");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("scatter_tma launch: ") + cudaGetErrorString(e));
}

void launch_scatter_generic(GenericParams p, int grid, cudaStream_t stream) {
  if (p.n_local == 0) return;
  const size_t budget = 160 * 1024;
  if (p.write_hi <= p.write_lo || p.write_hi > p.row_pitch || (p.write_lo & 3) || (p.write_hi & 3))
    throw std::runtime_error("scatter_generic: bad write range");
  const uint32_t span = p.write_hi - p.write_lo;
  uint32_t rows = static_cast<uint32_t>(budget / (span + 4 + sizeof(unsigned long long)));
  // 128 rows per block keeps the staging tile small enough for several CTAs per
  // SM: the per-thread field loop is latency bound, occupancy hides it.
  rows = rows >= 128 ? 128 : (rows / 32) * 32;
  if (rows == 0) throw std::runtime_error("row pitch too large for the generic scatter kernel");
  p.rows_per_block = rows;
  const size_t smem = static_cast<size_t>(rows) * (span + 4 + sizeof(unsigned long long)) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(scatter_generic_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(budget + 4096));
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured = budget + 4096;
  }
  if (grid <= 0) {
    int per_sm = static_cast<int>(std::min<size_t>(8, (200 * 1024) / (smem + 1024)));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long blocks = (p.n_local + rows - 1) / rows;
    grid = static_cast<int>(std::min<unsigned long long>(blocks, static_cast<unsigned long long>(sms) * std::max(1, per_sm)));
  }
  scatter_generic_kernel<<<grid, 256, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("scatter_generic launch: ") + cudaGetErrorString(e));
}

void launch_scatter_wide(const WideParams& p, int grid, cudaStream_t stream) {
  if (p.n_local == 0) return;
  if (grid <= 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long blocks = (p.n_local + 7) / 8;
    grid = static_cast<int>(std::min<unsigned long long>(blocks, static_cast<unsigned long long>(sms) * 8));
  }
  scatter_wide_kernel<<<grid, 256, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("scatter_wide launch: ") + cudaGetErrorString(e));
}

static void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

void launch_perm_positions(const PermKeyDev& key, const PlanDev& plan, unsigned long long global_offset,
                           unsigned long long n_local, int32_t* trainer, long long* slot,
                           cudaStream_t stream) {
  if (n_local == 0) return;
  int grid = static_cast<int>(std::min<unsigned long long>((n_local + 255) / 256, 148 * 8));
  perm_positions_kernel<<<grid, 256, 0, stream>>>(key, plan, global_offset, n_local, trainer, slot);
  check_launch("perm_positions");
}

void launch_place_rows(const uint8_t* rows, const long long* src_idx, const long long* dst_off,
                       unsigned long long n, uint32_t pitch, uint8_t* dst, cudaStream_t stream) {
  if (n == 0) return;
  if (pitch % 16) throw std::runtime_error("place_rows: row pitch must be a multiple of 16");
  int grid = static_cast<int>(std::min<unsigned long long>((n + 7) / 8, 148 * 16));
  place_rows_kernel<<<grid, 256, 0, stream>>>(rows, src_idx, dst_off, n, pitch, dst);
  check_launch("place_rows");
}

void launch_key_checksum(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                         uint32_t key_off, unsigned long long* out, cudaStream_t stream) {
  if (rows == 0) return;
  int grid = static_cast<int>(std::min<unsigned long long>((rows + 255) / 256, 148 * 4));
  key_checksum_kernel<<<grid, 256, 0, stream>>>(packed, rows, pitch, key_off, out);
  check_launch("key_checksum");
}

void launch_batch_sum_f32(const uint8_t* packed, unsigned long long rows, uint32_t pitch,
                          uint32_t off, double* out, cudaStream_t stream) {
  if (rows == 0) return;
  int grid = static_cast<int>(std::min<unsigned long long>((rows + 255) / 256, 148 * 4));
  batch_sum_f32_kernel<<<grid, 256, 0, stream>>>(packed, rows, pitch, off, out);
  check_launch("batch_sum_f32");
}

void launch_batch_sum_all_f32(const uint8_t* packed, unsigned long long nbytes, double* out,
                              cudaStream_t stream) {
  const unsigned long long nvec = nbytes / 16;
  if (nvec == 0) return;
  // CTAs per SM x threads per CTA of the bench sink (A/B knob: RSDL_SINK_GRID="ctas,threads")
  static int ctas = 0, threads = 0;
  if (ctas == 0) {
    ctas = 4; threads = 256;
    if (const char* e = std::getenv("RSDL_SINK_GRID")) {
      int c = 0, t = 0;
      if (std::sscanf(e, "%d,%d", &c, &t) == 2 && c > 0 && c <= 16 && t >= 32 && t <= 256 && t % 32 == 0) {
        ctas = c; threads = t;
      }
    }
  }
  int grid = static_cast<int>(std::min<unsigned long long>((nvec + threads - 1) / threads, 148ull * ctas));
  batch_sum_all_f32_kernel<<<grid, threads, 0, stream>>>(reinterpret_cast<const float4*>(packed), nvec, out);
  check_launch("batch_sum_all_f32");
}

void launch_signal_flags(const FlagTargets& t, uint32_t value, cudaStream_t stream) {
  if (t.count == 0) return;
  signal_flags_kernel<<<1, RSDL_MAX_TRAINERS, 0, stream>>>(t, value);
  check_launch("signal_flags");
}

void launch_wait_flags(const uint32_t* flags, uint32_t count, uint32_t value,
                       unsigned long long timeout_ns, uint32_t* error, cudaStream_t stream) {
  if (count == 0) return;
  wait_flags_kernel<<<1, RSDL_MAX_TRAINERS, 0, stream>>>(flags, count, value, timeout_ns, error);
  check_launch("wait_flags");
}

}  // namespace rsdl
