// Python bindings + host runtime for the B200 shuffle engine.
//
// This is the native replacement for the roles the reference delegates to Ray
// (SURVEY.md 2.2): HBM arenas and pinned staging (plasma object store), CUDA-IPC
// peer mapping (object manager / named-actor directory), streams + events
// (task scheduler), epoch-tagged signal words (actor RPC), and a host thread
// pool for staging copies. Only the CUDA runtime is linked (statically); no
// libtorch dependency, tensors are exchanged as raw device pointers.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include <algorithm>
#include <cctype>
#include <fstream>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "kernels.h"
#include "perm.cuh"

namespace py = pybind11;
using rsdl::FastParams;
using rsdl::FieldDev;
using rsdl::FlagTargets;
using rsdl::GenericParams;
using rsdl::WideParams;

namespace {

inline void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T> T* as_ptr(uintptr_t v) { return reinterpret_cast<T*>(v); }
inline cudaStream_t as_stream(uintptr_t v) { return reinterpret_cast<cudaStream_t>(v); }
inline cudaEvent_t as_event(uintptr_t v) { return reinterpret_cast<cudaEvent_t>(v); }

PermKeyDev make_key(const std::vector<uint64_t>& w) {
  // (n, bits_l, bits_r, k0..k5) - ops/perm.py::PermKey.as_words
  if (w.size() != 3 + RSDL_PERM_ROUNDS) throw std::runtime_error("bad permutation key");
  PermKeyDev k;
  k.n = w[0];
  const uint32_t bits_l = static_cast<uint32_t>(w[1]);
  k.bits_r = static_cast<uint32_t>(w[2]);
  k.mask_l = bits_l >= 32 ? 0xFFFFFFFFu : ((1u << bits_l) - 1u);
  k.mask_r = k.bits_r >= 32 ? 0xFFFFFFFFu : ((1u << k.bits_r) - 1u);
  for (int i = 0; i < RSDL_PERM_ROUNDS; ++i) k.k[i] = static_cast<uint32_t>(w[3 + i]);
  return k;
}

PlanDev make_plan(uint64_t num_rows, uint32_t num_trainers, uint64_t slot_lo = 0,
          uint64_t slot_hi = ~0ull) {
  PlanDev p;
  p.slot_lo = slot_lo;
  p.slot_hi = slot_hi;
  p.q = num_rows / num_trainers;
  p.rem = static_cast<uint32_t>(num_rows % num_trainers);
  p.big = static_cast<unsigned long long>(p.rem) * (p.q + 1);
  p.num_trainers = num_trainers;
  return p;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from this driver");
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D view [num_cols][rows_alloc] of 4-byte source columns with a uniform
// stride; boxes are [panel cols][32 rows] = 128-byte rows, 128-byte swizzle.
void make_source_tmap(CUtensorMap* map, uintptr_t base, uint64_t rows_alloc, uint32_t num_cols,
                      uint64_t col_stride_bytes, uint32_t panel_cols, uint32_t box_rows,
                      bool swizzle) {
  cuuint64_t dims[2] = {rows_alloc, num_cols};
  cuuint64_t strides[1] = {col_stride_bytes};
  cuuint32_t box[2] = {box_rows, panel_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
                                 reinterpret_cast<void*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string(r));
}

template <typename P>
void fill_dst(P& p, const std::vector<uintptr_t>& dst) {
  if (dst.size() > RSDL_MAX_TRAINERS) throw std::runtime_error("too many trainers");
  for (size_t i = 0; i < RSDL_MAX_TRAINERS; ++i)
    p.dst[i] = i < dst.size() ? as_ptr<uint8_t>(dst[i]) : nullptr;
}

// ---------------------------------------------------------------------------
// Host worker pool: parallel staging copies (pageable -> pinned) off the GIL.
// ---------------------------------------------------------------------------
class HostPool {
 public:
  explicit HostPool(int n) : stop_(false) {
    for (int i = 0; i < std::max(1, n); ++i) workers_.emplace_back([this] { run(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return static_cast<int>(workers_.size()); }

  // Run fn(begin, end) over [0, n) in `grain`-sized pieces on the workers; wait.
  void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn) {
    if (n == 0) return;
    grain = std::max<size_t>(1, grain);
    const size_t pieces = (n + grain - 1) / grain;
    std::atomic<size_t> pending{pieces};
    std::mutex done_mu;
    std::condition_variable done_cv;
    for (size_t b = 0; b < n; b += grain) {
      const size_t e = std::min(n, b + grain);
      submit([=, &fn, &pending, &done_mu, &done_cv] {
        fn(b, e);
        if (pending.fetch_sub(1) == 1) {
          std::lock_guard<std::mutex> g(done_mu);
          done_cv.notify_all();
        }
      });
    }
    std::unique_lock<std::mutex> lk(done_mu);
    done_cv.wait(lk, [&] { return pending.load() == 0; });
  }

  // Split [0, nbytes) into chunks, copy them on the workers, wait for all.
  void parallel_memcpy(void* dst, const void* src, size_t nbytes) {
    const size_t chunk = std::max<size_t>(1 << 20, nbytes / (workers_.size() * 4) + 1);
    std::atomic<size_t> pending{0};
    std::mutex done_mu;
    std::condition_variable done_cv;
    size_t launched = 0;
    for (size_t off = 0; off < nbytes; off += chunk) ++launched;
    pending = launched;
    for (size_t off = 0; off < nbytes; off += chunk) {
      const size_t len = std::min(chunk, nbytes - off);
      submit([=, &pending, &done_mu, &done_cv] {
        std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len);
        if (pending.fetch_sub(1) == 1) {
          std::lock_guard<std::mutex> g(done_mu);
          done_cv.notify_all();
        }
      });
    }
    std::unique_lock<std::mutex> lk(done_mu);
    done_cv.wait(lk, [&] { return pending.load() == 0; });
  }

 private:
  void submit(std::function<void()> fn) {
    {
      std::lock_guard<std::mutex> g(mu_);
      tasks_.push(std::move(fn));
    }
    cv_.notify_one();
  }
  void run() {
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !tasks_.empty(); });
        if (stop_ && tasks_.empty()) return;
        fn = std::move(tasks_.front());
        tasks_.pop();
      }
      fn();
    }
  }
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> tasks_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_;
};

// ---------------------------------------------------------------------------
// Flag polling on the host: the blocking half of the device ring (K11).
// ---------------------------------------------------------------------------
class FlagPoller {
 public:
  FlagPoller() {
    check(cudaHostAlloc(reinterpret_cast<void**>(&scratch_), RSDL_MAX_TRAINERS * sizeof(uint32_t),
                        cudaHostAllocDefault), "cudaHostAlloc(flag scratch)");
    check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate(flag)");
  }
  ~FlagPoller() {
    cudaStreamDestroy(stream_);
    cudaFreeHost(scratch_);
  }
  // Returns the index of the first flag still below `value` after `timeout_s`
  // seconds, or -1 once all `count` flags are >= value (wrap-safe compare).
  int wait(const uint32_t* dev_flags, uint32_t count, uint32_t value, double timeout_s) {
    if (count > RSDL_MAX_TRAINERS) throw std::runtime_error("too many flags");
    const auto t0 = std::chrono::steady_clock::now();
    int sleep_us = 5;
    for (;;) {
      int lagging = -1;
      {
        // several consumer threads may poll through one engine: the scratch
        // buffer and the stream are shared
        std::lock_guard<std::mutex> g(mu_);
        check(cudaMemcpyAsync(scratch_, dev_flags, count * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, stream_), "cudaMemcpyAsync(flags)");
        check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize(flags)");
        for (uint32_t i = 0; i < count; ++i)
          if (static_cast<int32_t>(scratch_[i] - value) < 0) { lagging = static_cast<int>(i); break; }
      }
      if (lagging < 0) return -1;
      const double waited =
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (timeout_s >= 0 && waited > timeout_s) return lagging;
      std::this_thread::sleep_for(std::chrono::microseconds(sleep_us));
      if (sleep_us < 200) sleep_us *= 2;
    }
  }
  std::vector<uint32_t> read(const uint32_t* dev_flags, uint32_t count) {
    std::lock_guard<std::mutex> g(mu_);
    check(cudaMemcpyAsync(scratch_, dev_flags, count * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          stream_), "cudaMemcpyAsync(flags)");
    check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize(flags)");
    return std::vector<uint32_t>(scratch_, scratch_ + count);
  }

 private:
  uint32_t* scratch_ = nullptr;
  cudaStream_t stream_ = nullptr;
  std::mutex mu_;
};

// ---------------------------------------------------------------------------
// Host cast + pack: columnar table -> packed rows (the CPU backend's analogue
// of the scatter kernels' epilogue; replaces one strided numpy assignment per
// column, reference torch_dataset.py:204-236). Cast rules mirror
// ops/layout.py::cast_column and the device's load_src / store_cast.
// ---------------------------------------------------------------------------
struct bf16_bits { uint16_t b; };
struct f16_val { _Float16 v; };
struct bool_val { uint8_t v; };

template <typename S> struct Wide { using type = long long; };
template <> struct Wide<float> { using type = float; };
template <> struct Wide<double> { using type = double; };
template <> struct Wide<bf16_bits> { using type = float; };
template <> struct Wide<f16_val> { using type = float; };

template <typename S> inline typename Wide<S>::type load_wide(const S& x) { return static_cast<typename Wide<S>::type>(x); }
template <> inline float load_wide<bf16_bits>(const bf16_bits& x) {
  uint32_t u = static_cast<uint32_t>(x.b) << 16; float f; std::memcpy(&f, &u, 4); return f; }
template <> inline float load_wide<f16_val>(const f16_val& x) { return static_cast<float>(x.v); }
template <> inline long long load_wide<bool_val>(const bool_val& x) { return x.v ? 1 : 0; }

template <typename D, typename W> struct Conv {
  static inline D run(W w) {
    if constexpr (std::is_integral<D>::value && !std::is_integral<W>::value)
      return static_cast<D>(static_cast<long long>(w));      // float -> int: via int64, then wrap
    else
      return static_cast<D>(w);
  }
};
template <typename W> struct Conv<bf16_bits, W> {
  static inline bf16_bits run(W w) {
    const float f = static_cast<float>(w);
    uint32_t u; std::memcpy(&u, &f, 4);
    if (f != f) return bf16_bits{0x7FFF};                    // NaN canonicalised like CUDA
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;              // round to nearest even
    return bf16_bits{static_cast<uint16_t>(u)};
  }
};
template <typename W> struct Conv<f16_val, W> {
  static inline f16_val run(W w) { return f16_val{static_cast<_Float16>(w)}; }
};
template <typename W> struct Conv<bool_val, W> {
  static inline bool_val run(W w) { return bool_val{static_cast<uint8_t>(w != 0 ? 1 : 0)}; }
};

template <typename S, typename D>
void pack_field_block(const uint8_t* src_base, uint32_t width, uint8_t* out, uint32_t pitch,
                      uint32_t dst_off, size_t r0, size_t r1) {
  const S* src = reinterpret_cast<const S*>(src_base);
  using W = typename Wide<S>::type;
  for (size_t r = r0; r < r1; ++r) {
    uint8_t* o = out + r * pitch + dst_off;
    for (uint32_t w = 0; w < width; ++w) {
      const D d = Conv<D, W>::run(load_wide<S>(src[r * width + w]));
      std::memcpy(o + w * sizeof(D), &d, sizeof(D));         // fields are only dst-size aligned
    }
  }
}

template <typename S>
bool pack_field_dst(uint32_t dst_code, const uint8_t* src, uint32_t width, uint8_t* out,
                    uint32_t pitch, uint32_t off, size_t r0, size_t r1) {
  switch (dst_code) {
    case DT_U8:   pack_field_block<S, uint8_t>(src, width, out, pitch, off, r0, r1); return true;
    case DT_I8:   pack_field_block<S, int8_t>(src, width, out, pitch, off, r0, r1); return true;
    case DT_I16:  pack_field_block<S, int16_t>(src, width, out, pitch, off, r0, r1); return true;
    case DT_I32:  pack_field_block<S, int32_t>(src, width, out, pitch, off, r0, r1); return true;
    case DT_I64:  pack_field_block<S, long long>(src, width, out, pitch, off, r0, r1); return true;
    case DT_F16:  pack_field_block<S, f16_val>(src, width, out, pitch, off, r0, r1); return true;
    case DT_BF16: pack_field_block<S, bf16_bits>(src, width, out, pitch, off, r0, r1); return true;
    case DT_F32:  pack_field_block<S, float>(src, width, out, pitch, off, r0, r1); return true;
    case DT_F64:  pack_field_block<S, double>(src, width, out, pitch, off, r0, r1); return true;
    case DT_BOOL: pack_field_block<S, bool_val>(src, width, out, pitch, off, r0, r1); return true;
    default: return false;                                    // fp8: numpy golden path
  }
}

bool pack_field(const FieldDev& f, uint8_t* out, uint32_t pitch, size_t r0, size_t r1) {
  const uint8_t* s = f.src;
  switch (f.src_code) {
    case DT_U8:   return pack_field_dst<uint8_t>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_I8:   return pack_field_dst<int8_t>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_I16:  return pack_field_dst<int16_t>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_I32:  return pack_field_dst<int32_t>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_I64:  return pack_field_dst<long long>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_F16:  return pack_field_dst<f16_val>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_BF16: return pack_field_dst<bf16_bits>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_F32:  return pack_field_dst<float>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_F64:  return pack_field_dst<double>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    case DT_BOOL: return pack_field_dst<bool_val>(f.dst_code, s, f.width, out, pitch, f.dst_off, r0, r1);
    default: return false;
  }
}

// ---------------------------------------------------------------------------
// NUMA placement: a rank's pinned staging memory and its host threads should
// sit on the socket its GPU hangs off (HGX B200: GPUs 0-3 on node 0, 4-7 on
// node 1) - with 8 ranks streaming from host memory at once, remote-socket
// pinned buffers cost ~15% of the PCIe rate (profiles/README.md, e2e at N=8).
// Ray's raylet does the equivalent per-node worker placement for the reference.
// ---------------------------------------------------------------------------
int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
  std::string id(bus);
  std::transform(id.begin(), id.end(), id.begin(), [](unsigned char c) { return std::tolower(c); });
  std::ifstream f("/sys/bus/pci/devices/" + id + "/numa_node");
  int node = -1;
  if (!(f >> node)) return -1;
  return node;
}

// "0-31,64-95" -> cpu ids
std::vector<int> parse_cpulist(const std::string& text) {
  std::vector<int> cpus;
  size_t i = 0;
  while (i < text.size()) {
    while (i < text.size() && !std::isdigit(static_cast<unsigned char>(text[i]))) ++i;
    if (i >= text.size()) break;
    int a = 0;
    while (i < text.size() && std::isdigit(static_cast<unsigned char>(text[i]))) a = a * 10 + (text[i++] - '0');
    int b = a;
    if (i < text.size() && text[i] == '-') {
      ++i;
      b = 0;
      while (i < text.size() && std::isdigit(static_cast<unsigned char>(text[i]))) b = b * 10 + (text[i++] - '0');
    }
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) cpus.push_back(c);
  }
  return cpus;
}

std::vector<int> numa_node_cpus(int node) {
  std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string line;
  if (!std::getline(f, line)) return {};
  return parse_cpulist(line);
}

// Restrict the calling thread (and every thread it creates afterwards) to the
// node's CPUs, intersected with the CPUs it is currently allowed to use, and
// prefer the node's memory for its allocations (first touch + MPOL_PREFERRED).
// Returns the number of CPUs in the new mask, 0 when nothing was changed.
int bind_thread_to_numa_node(int node) {
  if (node < 0) return 0;
  const std::vector<int> cpus = numa_node_cpus(node);
  if (cpus.empty()) return 0;
  cpu_set_t allowed, want;
  CPU_ZERO(&allowed);
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return 0;
  int n = 0;
  for (int c : cpus)
    if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ++n; }
  if (n == 0) return 0;                       // cgroup excludes that socket: leave as is
  if (sched_setaffinity(0, sizeof(want), &want) != 0) return 0;
#ifdef SYS_set_mempolicy
  if (node < 64) {
    unsigned long mask = 1ul << node;
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8 + 1);
  }
#endif
  return n;
}

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.doc() = "ray_shuffling_data_loader_b200 native runtime (sm_100a)";
  m.attr("MAX_TRAINERS") = RSDL_MAX_TRAINERS;
  m.attr("TILE_ROWS") = rsdl::fast_max_tile_rows();   // source-column padding granule
  m.def("fast_tile_rows", &rsdl::fast_tile_rows);
  m.def("fast_panel_cols", &rsdl::fast_panel_cols);

  // ---- device -----------------------------------------------------------
  m.def("device_count", [] { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; });
  m.def("set_device", [](int d) { check(cudaSetDevice(d), "cudaSetDevice"); });
  m.def("get_device", [] { int d; check(cudaGetDevice(&d), "cudaGetDevice"); return d; });
  m.def("sm_count", [](int d) {
    int v; check(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d), "attr"); return v; });
  m.def("compute_capability", [](int d) {
    int a, b;
    check(cudaDeviceGetAttribute(&a, cudaDevAttrComputeCapabilityMajor, d), "attr");
    check(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMinor, d), "attr");
    return std::make_pair(a, b); });
  m.def("gpu_numa_node", &gpu_numa_node,
        "NUMA node of the GPU's PCIe root (-1 when the platform does not say)");
  m.def("numa_node_cpus", &numa_node_cpus);
  m.def("parse_cpulist", &parse_cpulist);
  m.def("bind_thread_to_numa_node", &bind_thread_to_numa_node,
        "pin the calling thread (and its future children) + memory policy to a NUMA node");
  m.def("mem_get_info", [] {
    size_t f, t; check(cudaMemGetInfo(&f, &t), "cudaMemGetInfo"); return std::make_pair(f, t); });
  m.def("can_access_peer", [](int a, int b) {
    int v = 0; check(cudaDeviceCanAccessPeer(&v, a, b), "cudaDeviceCanAccessPeer"); return v != 0; });
  m.def("device_synchronize", [] {
    py::gil_scoped_release r; check(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); });

  // ---- memory: HBM arenas, pinned staging, CUDA IPC --------------------------
  m.def("device_malloc", [](size_t n) {
    void* p = nullptr; check(cudaMalloc(&p, std::max<size_t>(n, 256)), "cudaMalloc");
    return reinterpret_cast<uintptr_t>(p); });
  m.def("device_free", [](uintptr_t p) { check(cudaFree(as_ptr<void>(p)), "cudaFree"); });
  m.def("device_memset_async", [](uintptr_t p, int v, size_t n, uintptr_t s) {
    check(cudaMemsetAsync(as_ptr<void>(p), v, n, as_stream(s)), "cudaMemsetAsync"); });
  m.def("pinned_alloc", [](size_t n) {
    void* p = nullptr;
    py::gil_scoped_release r;
    check(cudaHostAlloc(&p, std::max<size_t>(n, 64), cudaHostAllocDefault), "cudaHostAlloc");
    return reinterpret_cast<uintptr_t>(p); });
  m.def("pinned_free", [](uintptr_t p) { check(cudaFreeHost(as_ptr<void>(p)), "cudaFreeHost"); });
  m.def("host_register", [](uintptr_t p, size_t n) {
    check(cudaHostRegister(as_ptr<void>(p), n, cudaHostRegisterDefault), "cudaHostRegister"); });
  m.def("host_unregister", [](uintptr_t p) {
    check(cudaHostUnregister(as_ptr<void>(p)), "cudaHostUnregister"); });
  m.def("ipc_get_handle", [](uintptr_t p) {
    cudaIpcMemHandle_t h;
    check(cudaIpcGetMemHandle(&h, as_ptr<void>(p)), "cudaIpcGetMemHandle");
    return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)); });
  m.def("ipc_open_handle", [](const std::string& b) {
    if (b.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("bad IPC handle");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, b.data(), sizeof(h));
    void* p = nullptr;
    check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return reinterpret_cast<uintptr_t>(p); });
  m.def("ipc_close_handle", [](uintptr_t p) {
    check(cudaIpcCloseMemHandle(as_ptr<void>(p)), "cudaIpcCloseMemHandle"); });
  m.def("enable_peer_access", [](int peer) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return; }
    check(e, "cudaDeviceEnablePeerAccess"); });

  // ---- copies ------------------------------------------------------------
  m.def("memcpy_async", [](uintptr_t dst, uintptr_t src, size_t n, int kind, uintptr_t s) {
    check(cudaMemcpyAsync(as_ptr<void>(dst), as_ptr<void>(src), n,
                          static_cast<cudaMemcpyKind>(kind), as_stream(s)), "cudaMemcpyAsync"); });
  m.def("memcpy2d_async", [](uintptr_t dst, size_t dpitch, uintptr_t src, size_t spitch,
                             size_t width, size_t height, int kind, uintptr_t s) {
    check(cudaMemcpy2DAsync(as_ptr<void>(dst), dpitch, as_ptr<void>(src), spitch, width, height,
                            static_cast<cudaMemcpyKind>(kind), as_stream(s)), "cudaMemcpy2DAsync"); });
  m.attr("H2D") = static_cast<int>(cudaMemcpyHostToDevice);
  m.attr("D2H") = static_cast<int>(cudaMemcpyDeviceToHost);
  m.attr("D2D") = static_cast<int>(cudaMemcpyDeviceToDevice);

  // ---- streams & events ----------------------------------------------------
  m.def("stream_priority_range", [] {
    int lo, hi; check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "priority range");
    return std::make_pair(lo, hi); });
  m.def("stream_create", [](int priority) {
    cudaStream_t s;
    check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority), "cudaStreamCreate");
    return reinterpret_cast<uintptr_t>(s); });
  m.def("stream_destroy", [](uintptr_t s) { check(cudaStreamDestroy(as_stream(s)), "cudaStreamDestroy"); });
  m.def("stream_synchronize", [](uintptr_t s) {
    py::gil_scoped_release r; check(cudaStreamSynchronize(as_stream(s)), "cudaStreamSynchronize"); });
  m.def("stream_query", [](uintptr_t s) {
    cudaError_t e = cudaStreamQuery(as_stream(s));
    if (e == cudaErrorNotReady) return false;
    check(e, "cudaStreamQuery"); return true; });
  m.def("stream_wait_event", [](uintptr_t s, uintptr_t e) {
    check(cudaStreamWaitEvent(as_stream(s), as_event(e), 0), "cudaStreamWaitEvent"); });
  m.def("event_create", [](bool timing) {
    cudaEvent_t e;
    check(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming),
          "cudaEventCreate");
    return reinterpret_cast<uintptr_t>(e); });
  m.def("event_destroy", [](uintptr_t e) { check(cudaEventDestroy(as_event(e)), "cudaEventDestroy"); });
  m.def("event_record", [](uintptr_t e, uintptr_t s) {
    check(cudaEventRecord(as_event(e), as_stream(s)), "cudaEventRecord"); });
  m.def("event_synchronize", [](uintptr_t e) {
    py::gil_scoped_release r; check(cudaEventSynchronize(as_event(e)), "cudaEventSynchronize"); });
  m.def("event_query", [](uintptr_t e) {
    cudaError_t r = cudaEventQuery(as_event(e));
    if (r == cudaErrorNotReady) return false;
    check(r, "cudaEventQuery"); return true; });
  m.def("event_elapsed_ms", [](uintptr_t a, uintptr_t b) {
    float ms = 0; check(cudaEventElapsedTime(&ms, as_event(a), as_event(b)), "cudaEventElapsedTime");
    return ms; });

  // ---- kernels ---------------------------------------------------------------
  m.def("scatter_fast",
        [](const std::vector<uint64_t>& key, uint64_t num_rows, uint32_t num_trainers,
           uintptr_t cols, uint32_t num_cols, uint64_t n_local, uint64_t global_offset,
           uint32_t row_pitch, uint32_t scale_offset, const std::vector<uintptr_t>& dst, int mode,
           int grid, uintptr_t stream, uintptr_t col_base, uint64_t col_stride,
           uint64_t rows_alloc, int tmap_mode, uintptr_t kinds, uint32_t write_end, int sched,
           uint64_t slot_lo, uint64_t slot_hi,
           const std::vector<std::tuple<uintptr_t, uint32_t, uint32_t, uint32_t>>& tail,
           uint32_t tail_lo, uint32_t tail_hi) {
          FastParams p;
          if (tail.size() > RSDL_MAX_TAIL_FIELDS)
            throw std::runtime_error("scatter_fast: too many tail fields");
          if (tail_hi > tail_lo && ((tail_lo | tail_hi) & 15u || tail_hi > row_pitch))
            throw std::runtime_error("scatter_fast: tail range must be 16-byte multiples inside the row");
          p.tail_lo = tail_lo;
          p.tail_hi = tail_hi > tail_lo ? tail_hi : tail_lo;
          p.num_tail = static_cast<uint32_t>(tail.size());
          p.pad_ = 0;
          for (size_t i = 0; i < RSDL_MAX_TAIL_FIELDS; ++i) {
            p.tail[i] = rsdl::TailField{nullptr, 0, 0, 0, 0};
            if (i < tail.size()) {
              const uint32_t dsz = rsdl_itemsize(std::get<2>(tail[i]));
              const uint32_t doff = std::get<3>(tail[i]);
              if ((dsz != 4 && dsz != 8) || doff % dsz || doff < tail_lo || doff + dsz > tail_hi)
                throw std::runtime_error("scatter_fast: bad tail field");
              p.tail[i] = rsdl::TailField{as_ptr<const uint8_t>(std::get<0>(tail[i])),
                                          std::get<1>(tail[i]), std::get<2>(tail[i]), doff, 0};
            }
          }
          std::memset(&p.tmap, 0, sizeof(p.tmap));
          p.use_tmap = 0;
          const uint32_t tile_rows = static_cast<uint32_t>(rsdl::fast_tile_rows(mode));
          if (tmap_mode != 0 && col_base != 0 && col_stride % 16 == 0 &&
              rows_alloc % tile_rows == 0 && rsdl::fast_src_itemsize(mode) == 4) {
            const bool dense = tmap_mode == 2;
            make_source_tmap(&p.tmap, col_base, rows_alloc, num_cols, col_stride,
                             static_cast<uint32_t>(rsdl::fast_panel_cols(mode)),
                             dense ? tile_rows : 32u, !dense);
            p.use_tmap = dense ? 2 : 1;
          }
          p.key = make_key(key);
          p.plan = make_plan(num_rows, num_trainers, slot_lo, slot_hi);
          p.cols = as_ptr<const uint8_t* const>(cols);
          p.kinds = as_ptr<const uint8_t>(kinds);
          p.num_cols = num_cols;
          const uint32_t panel = static_cast<uint32_t>(rsdl::fast_panel_cols(mode));
          p.num_panels = (num_cols + panel - 1) / panel;
          p.n_local = n_local;
          p.global_offset = global_offset;
          p.row_pitch = row_pitch;
          p.scale_offset = scale_offset;
          if (write_end > row_pitch || (write_end & 15u))
            throw std::runtime_error("scatter_fast: write_end must be a 16-byte multiple <= row_pitch");
          p.write_end = write_end;
          // auto: the cooperative schedule only wins for ~256-byte f32 rows
          // (profiles/kbench_v17.jsonl: 64 cols 1.09 vs 1.20 ms; 32/48/96/128+ cols
          // equal or worse; bf16 and 8-byte sources worse)
          p.sched = sched >= 0 ? static_cast<uint32_t>(sched)
                               : ((mode == 0 && num_cols > 48 && num_cols <= 80) ? 1u : 0u);
          fill_dst(p, dst);
          rsdl::launch_scatter_fast(p, mode, grid, as_stream(stream));
        },
        py::arg("key"), py::arg("num_rows"), py::arg("num_trainers"), py::arg("cols"),
        py::arg("num_cols"), py::arg("n_local"), py::arg("global_offset"), py::arg("row_pitch"),
        py::arg("scale_offset"), py::arg("dst"), py::arg("mode"), py::arg("grid"),
        py::arg("stream"), py::arg("col_base") = 0, py::arg("col_stride") = 0,
        py::arg("rows_alloc") = 0, py::arg("tmap_mode") = 2, py::arg("kinds") = 0,
        py::arg("write_end") = 0, py::arg("sched") = -1, py::arg("slot_lo") = 0,
        py::arg("slot_hi") = ~0ull,
        py::arg("tail") = std::vector<std::tuple<uintptr_t, uint32_t, uint32_t, uint32_t>>{},
        py::arg("tail_lo") = 0, py::arg("tail_hi") = 0);
  m.def("fast_src_itemsize", &rsdl::fast_src_itemsize);
  m.def("fast_ctas_per_sm", &rsdl::fast_ctas_per_sm);
  m.def("scatter_generic",
        [](const std::vector<uint64_t>& key, uint64_t num_rows, uint32_t num_trainers,
           uintptr_t fields, uint32_t num_fields, uint64_t n_local, uint64_t global_offset,
           uint32_t row_pitch, uint32_t write_lo, uint32_t write_hi,
           const std::vector<uintptr_t>& dst, int grid, uintptr_t stream, uint64_t slot_lo,
           uint64_t slot_hi) {
          GenericParams p;
          p.key = make_key(key);
          p.plan = make_plan(num_rows, num_trainers, slot_lo, slot_hi);
          p.fields = as_ptr<const FieldDev>(fields);
          p.num_fields = num_fields;
          p.rows_per_block = 0;
          p.n_local = n_local;
          p.global_offset = global_offset;
          p.row_pitch = row_pitch;
          p.write_lo = write_lo;
          p.write_hi = write_hi;
          fill_dst(p, dst);
          rsdl::launch_scatter_generic(p, grid, as_stream(stream));
        },
        py::arg("key"), py::arg("num_rows"), py::arg("num_trainers"), py::arg("fields"),
        py::arg("num_fields"), py::arg("n_local"), py::arg("global_offset"),
        py::arg("row_pitch"), py::arg("write_lo"), py::arg("write_hi"), py::arg("dst"),
        py::arg("grid"), py::arg("stream"), py::arg("slot_lo") = 0, py::arg("slot_hi") = ~0ull);
  m.def("scatter_wide",
        [](const std::vector<uint64_t>& key, uint64_t num_rows, uint32_t num_trainers,
           uintptr_t src, uint32_t width, uint32_t src_code, uint32_t dst_code, uint32_t dst_off,
           uint64_t n_local, uint64_t global_offset, uint32_t row_pitch,
           const std::vector<uintptr_t>& dst, int grid, uintptr_t stream, uint64_t slot_lo,
           uint64_t slot_hi) {
          WideParams p;
          p.key = make_key(key);
          p.plan = make_plan(num_rows, num_trainers, slot_lo, slot_hi);
          p.src = as_ptr<const uint8_t>(src);
          p.width = width;
          p.src_code = src_code;
          p.dst_code = dst_code;
          p.dst_off = dst_off;
          p.n_local = n_local;
          p.global_offset = global_offset;
          p.row_pitch = row_pitch;
          fill_dst(p, dst);
          rsdl::launch_scatter_wide(p, grid, as_stream(stream));
        },
        py::arg("key"), py::arg("num_rows"), py::arg("num_trainers"), py::arg("src"),
        py::arg("width"), py::arg("src_code"), py::arg("dst_code"), py::arg("dst_off"),
        py::arg("n_local"), py::arg("global_offset"), py::arg("row_pitch"), py::arg("dst"),
        py::arg("grid"), py::arg("stream"), py::arg("slot_lo") = 0, py::arg("slot_hi") = ~0ull);
  m.attr("FIELD_DESC_BYTES") = sizeof(FieldDev);
  m.def("perm_positions",
        [](const std::vector<uint64_t>& key, uint64_t num_rows, uint32_t num_trainers,
           uint64_t global_offset, uint64_t n_local, uintptr_t trainer, uintptr_t slot,
           uintptr_t stream) {
          rsdl::launch_perm_positions(make_key(key), make_plan(num_rows, num_trainers),
                                      global_offset, n_local, as_ptr<int32_t>(trainer),
                                      as_ptr<long long>(slot), as_stream(stream));
        });
  m.def("place_rows", [](uintptr_t rows, uintptr_t dst_off, uint64_t n, uint32_t pitch,
                         uintptr_t dst, uintptr_t stream, uintptr_t src_idx) {
    rsdl::launch_place_rows(as_ptr<const uint8_t>(rows), as_ptr<const long long>(src_idx),
                            as_ptr<const long long>(dst_off), n, pitch,
                            as_ptr<uint8_t>(dst), as_stream(stream));
  }, py::arg("rows"), py::arg("dst_off"), py::arg("n"), py::arg("pitch"), py::arg("dst"),
     py::arg("stream"), py::arg("src_idx") = 0,
     "dst[dst_off[i] bytes] <- rows[src_idx[i] or i] (row copies of `pitch` bytes; "
     "negative offsets are skipped)");
  m.def("key_checksum", [](uintptr_t packed, uint64_t rows, uint32_t pitch, uint32_t key_off,
                           uintptr_t out, uintptr_t stream) {
    rsdl::launch_key_checksum(as_ptr<const uint8_t>(packed), rows, pitch, key_off,
                              as_ptr<unsigned long long>(out), as_stream(stream));
  });
  m.def("batch_sum_f32", [](uintptr_t packed, uint64_t rows, uint32_t pitch, uint32_t off,
                            uintptr_t out, uintptr_t stream) {
    rsdl::launch_batch_sum_f32(as_ptr<const uint8_t>(packed), rows, pitch, off,
                               as_ptr<double>(out), as_stream(stream));
  });
  m.def("batch_sum_all_f32", [](uintptr_t packed, uint64_t nbytes, uintptr_t out, uintptr_t stream) {
    rsdl::launch_batch_sum_all_f32(as_ptr<const uint8_t>(packed), nbytes, as_ptr<double>(out),
                                   as_stream(stream));
  });
  m.def("bulk_store_probe", [](uintptr_t dst, uint64_t slots, uint32_t row_bytes,
                               uint64_t total_rows, int grid, uintptr_t stream) {
    rsdl::launch_bulk_store_probe(as_ptr<uint8_t>(dst), slots, row_bytes, total_rows, grid,
                                  as_stream(stream));
  }, "rate probe: TMA bulk stores (smem -> global) of `row_bytes` rows to random slots");
  m.def("signal_flags", [](const std::vector<uintptr_t>& ptrs, uint32_t value, uintptr_t stream) {
    if (ptrs.size() > RSDL_MAX_TRAINERS) throw std::runtime_error("too many flag targets");
    FlagTargets t;
    t.count = static_cast<uint32_t>(ptrs.size());
    for (size_t i = 0; i < RSDL_MAX_TRAINERS; ++i)
      t.ptr[i] = i < ptrs.size() ? as_ptr<uint32_t>(ptrs[i]) : nullptr;
    rsdl::launch_signal_flags(t, value, as_stream(stream));
  });
  m.def("wait_flags", [](uintptr_t flags, uint32_t count, uint32_t value, uint64_t timeout_ns,
                         uintptr_t error, uintptr_t stream) {
    rsdl::launch_wait_flags(as_ptr<const uint32_t>(flags), count, value, timeout_ns,
                            as_ptr<uint32_t>(error), as_stream(stream));
  });

  // ---- host runtime objects --------------------------------------------------
  py::class_<HostPool>(m, "HostPool")
      .def(py::init<int>(), py::arg("num_threads"))
      .def_property_readonly("size", &HostPool::size)
      .def("parallel_memcpy",
           [](HostPool& self, uintptr_t dst, uintptr_t src, size_t n) {
             py::gil_scoped_release r;
             self.parallel_memcpy(as_ptr<void>(dst), as_ptr<const void>(src), n);
           });
  // ---- host shuffle (backend="cpu"): the same bijection, on the worker pool ----
  // The CPU engine's per-epoch work (reference shuffle_map + shuffle_reduce,
  // shuffle.py:129-200) as native code: perm.cuh is shared with the kernels, so
  // host and device agree bit for bit by construction.
  m.def("host_pack_rows",
        [](HostPool& pool, uintptr_t fields, uint32_t num_fields, uint64_t num_rows,
           uint32_t row_pitch, uintptr_t out) {
          // `fields`: host array of FieldDev (src = host column pointer)
          const FieldDev* f = as_ptr<const FieldDev>(fields);
          for (uint32_t i = 0; i < num_fields; ++i)
            if (f[i].dst_code == DT_FP8 || f[i].src_code == DT_FP8)
              throw std::runtime_error("host_pack_rows: fp8 fields take the numpy path");
          uint8_t* o = as_ptr<uint8_t>(out);
          std::atomic<bool> ok{true};
          py::gil_scoped_release r;
          pool.parallel_for(num_rows, 4096, [&](size_t b, size_t e) {
            std::memset(o + b * row_pitch, 0, (e - b) * static_cast<size_t>(row_pitch));
            for (uint32_t i = 0; i < num_fields; ++i)
              if (!pack_field(f[i], o, row_pitch, b, e)) ok = false;
          });
          if (!ok) throw std::runtime_error("host_pack_rows: unsupported dtype code");
        },
        py::arg("pool"), py::arg("fields"), py::arg("num_fields"), py::arg("num_rows"),
        py::arg("row_pitch"), py::arg("out"));
  // packed rows -> one contiguous array per field (the inverse of host_pack_rows without
  // casts: fields keep their packed dtype). `fields`: host array of FieldDev whose `src`
  // is the DESTINATION column base and whose `width * itemsize(dst_code)` bytes at
  // `dst_off` of every row are copied. Used to turn a reducer chunk into DataFrame
  // columns once, on all cores, instead of slicing rows batch by batch in Python.
  m.def("host_unpack_fields",
        [](HostPool& pool, uintptr_t packed, uint64_t num_rows, uint32_t row_pitch,
           uintptr_t fields, uint32_t num_fields) {
          const FieldDev* f = as_ptr<const FieldDev>(fields);
          const uint8_t* in = as_ptr<const uint8_t>(packed);
          py::gil_scoped_release r;
          pool.parallel_for(num_rows, 8192, [&](size_t b, size_t e) {
            for (uint32_t i = 0; i < num_fields; ++i) {
              const size_t nb = static_cast<size_t>(f[i].width) * rsdl_itemsize(f[i].dst_code);
              uint8_t* out = const_cast<uint8_t*>(f[i].src);
              const uint8_t* src = in + f[i].dst_off;
              if (nb == 8) {
                for (size_t r2 = b; r2 < e; ++r2)
                  std::memcpy(out + r2 * 8, src + r2 * row_pitch, 8);
              } else if (nb == 4) {
                for (size_t r2 = b; r2 < e; ++r2)
                  std::memcpy(out + r2 * 4, src + r2 * row_pitch, 4);
              } else {
                for (size_t r2 = b; r2 < e; ++r2)
                  std::memcpy(out + r2 * nb, src + r2 * row_pitch, nb);
              }
            }
          });
        },
        py::arg("pool"), py::arg("packed"), py::arg("num_rows"), py::arg("row_pitch"),
        py::arg("fields"), py::arg("num_fields"));
  m.def("host_perm_positions",
        [](HostPool& pool, const std::vector<uint64_t>& key, uint64_t num_rows,
           uint32_t num_trainers, uint64_t global_offset, uint64_t n_local, uintptr_t trainer,
           uintptr_t slot) {
          const PermKeyDev k = make_key(key);
          const PlanDev plan = make_plan(num_rows, num_trainers);
          int32_t* tr = as_ptr<int32_t>(trainer);
          long long* sl = as_ptr<long long>(slot);
          py::gil_scoped_release r;
          pool.parallel_for(n_local, 1 << 16, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
              uint32_t t; unsigned long long s;
              rsdl_position_to_dest(rsdl_permute(global_offset + i, k), plan, &t, &s);
              tr[i] = static_cast<int32_t>(t);
              sl[i] = static_cast<long long>(s);
            }
          });
        },
        py::arg("pool"), py::arg("key"), py::arg("num_rows"), py::arg("num_trainers"),
        py::arg("global_offset"), py::arg("n_local"), py::arg("trainer"), py::arg("slot"));
  m.def("host_scatter_rows",
        [](HostPool& pool, const std::vector<uint64_t>& key, uint64_t num_rows,
           uint32_t num_trainers, uintptr_t packed, uint32_t row_pitch, uint64_t global_offset,
           uint64_t n_local, const std::vector<uintptr_t>& dst) {
          if (dst.size() != num_trainers) throw std::runtime_error("one destination per trainer");
          const PermKeyDev k = make_key(key);
          const PlanDev plan = make_plan(num_rows, num_trainers);
          const uint8_t* src = as_ptr<const uint8_t>(packed);
          py::gil_scoped_release r;
          pool.parallel_for(n_local, 1 << 14, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
              uint32_t t; unsigned long long s;
              rsdl_position_to_dest(rsdl_permute(global_offset + i, k), plan, &t, &s);
              if (dst[t] == 0) continue;            // trainer not served by this process
              std::memcpy(as_ptr<uint8_t>(dst[t]) + s * row_pitch, src + i * row_pitch, row_pitch);
            }
          });
        },
        py::arg("pool"), py::arg("key"), py::arg("num_rows"), py::arg("num_trainers"),
        py::arg("packed"), py::arg("row_pitch"), py::arg("global_offset"), py::arg("n_local"),
        py::arg("dst"));
  // K7 on the host backend: deliver destination positions [pos_lo, pos_hi) in order by
  // *pulling* each one's source row through the inverse permutation. Same bytes as the
  // scatter above, but a destination chunk is complete the moment its range is done, so
  // the CPU engine can hand out reducer chunks one by one.
  m.def("host_gather_rows",
        [](HostPool& pool, const std::vector<uint64_t>& key, uintptr_t packed, uint32_t row_pitch,
           uint64_t global_offset, uint64_t n_local, uint64_t pos_lo, uint64_t pos_hi,
           uintptr_t dst) {
          const PermKeyDev k = make_key(key);
          const uint8_t* src = as_ptr<const uint8_t>(packed);
          uint8_t* out = as_ptr<uint8_t>(dst);
          py::gil_scoped_release r;
          pool.parallel_for(pos_hi - pos_lo, 1 << 14, [&](size_t b, size_t e) {
            // random row *reads* are latency bound: resolve a block of source rows and
            // prefetch their first lines before copying any of them
            constexpr size_t kBlock = 32;
            unsigned long long gi[kBlock];
            for (size_t i0 = b; i0 < e; i0 += kBlock) {
              const size_t m = std::min(kBlock, e - i0);
              for (size_t j = 0; j < m; ++j) {
                gi[j] = rsdl_permute_inv(pos_lo + i0 + j, k);
                if (gi[j] >= global_offset && gi[j] < global_offset + n_local) {
                  const uint8_t* row = src + (gi[j] - global_offset) * row_pitch;
                  for (uint32_t o = 0; o < row_pitch; o += 64) __builtin_prefetch(row + o, 0, 0);
                }
              }
              for (size_t j = 0; j < m; ++j) {
                if (gi[j] < global_offset || gi[j] >= global_offset + n_local)
                  continue;                           // row owned by another process
                std::memcpy(out + (i0 + j) * row_pitch,
                            src + (gi[j] - global_offset) * row_pitch, row_pitch);
              }
            }
          });
        },
        py::arg("pool"), py::arg("key"), py::arg("packed"), py::arg("row_pitch"),
        py::arg("global_offset"), py::arg("n_local"), py::arg("pos_lo"), py::arg("pos_hi"),
        py::arg("dst"));
  py::class_<FlagPoller>(m, "FlagPoller")
      .def(py::init<>())
      .def("wait",
           [](FlagPoller& self, uintptr_t flags, uint32_t count, uint32_t value, double timeout_s) {
             py::gil_scoped_release r;
             return self.wait(as_ptr<const uint32_t>(flags), count, value, timeout_s);
           })
      .def("read", [](FlagPoller& self, uintptr_t flags, uint32_t count) {
        return self.read(as_ptr<const uint32_t>(flags), count);
      });
}
