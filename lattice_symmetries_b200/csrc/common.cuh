// common.cuh -- shared plumbing for liblattice_symmetries_b200.so
//
// One process drives one device (one rank per GPU).  All library work is
// ordered on a single non-blocking CUDA stream; host entry points are
// serialised by a mutex because the reference's callers invoke the per-state
// kernels concurrently from many tasks on the same private_data
// (chapel/src/StatesEnumeration.chpl:415-421).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/lattice_symmetries_b200.h"

namespace lsb {

// ---- error reporting (kernels/reference.c:11-36 semantics) -----------------
[[noreturn]] inline void fatal(char const *func, int line, char const *msg) {
  ls_hs_fatal_error(func, line, msg);
  abort();
}
#define LSB_CHECK(cond, msg) ((cond) ? (void)0 : ::lsb::fatal(__func__, __LINE__, msg))

// CUDA failures are unrecoverable for the call; they go through ls_hs_error so
// that a host-installed handler (Python: RuntimeError) sees them.
struct CudaFailure {
  std::string what;
};
inline void cuda_check(cudaError_t e, char const *expr, char const *file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e),
             cudaGetErrorString(e), file, line, expr);
    throw CudaFailure{buf};
  }
}
#define CUDA_CHECK(expr) ::lsb::cuda_check((expr), #expr, __FILE__, __LINE__)

// ---- process-wide runtime ---------------------------------------------------
struct Runtime {
  std::mutex mutex;  // serialises host entry points
  cudaStream_t stream = nullptr;      // the stream work is ordered on (own_stream unless overridden)
  cudaStream_t own_stream = nullptr;
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  std::atomic<uint64_t> launches{0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_matvec_ms = 0, last_build_ms = 0;
  // LS_B200_PROFILE=1: summed device time / launch count of the two matvec kernels in the last matvec
  double last_orbit_ms = 0, last_gather_ms = 0, last_combine_ms = 0, last_allgather_ms = 0, last_count_ms = 0;
  int last_orbit_launches = 0, last_gather_launches = 0;

  void ensure();  // throws CudaFailure when no usable device
};
Runtime &runtime();

// block cache (runtime.cu): large device allocations are recycled instead of going back to the driver
void *block_alloc(size_t bytes, bool managed = false);
void block_free(void *p);            // any device pointer
size_t block_cache_trim(size_t keep_bytes = 0);
size_t block_cache_idle_bytes();
template <class T>
inline void block_alloc(T **p, size_t bytes, bool managed = false) { *p = static_cast<T *>(block_alloc(bytes, managed)); }
void *alloc_local(size_t bytes);  // large randomly-accessed tables (runtime.cu)
template <class T>
inline void alloc_local(T **p, size_t bytes) { *p = static_cast<T *>(alloc_local(bytes)); }
inline void count_launch(int n = 1) { runtime().launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Wraps an extern "C" entry point body: serialise, translate failures.
// The library owns ONE device per process (one rank per GPU); a caller may have another device current on the
// calling thread (torch.cuda.device(...), a different host thread), so every entry point switches to the
// library's device for its duration and restores the caller's choice afterwards.
struct DeviceScope {
  int previous = -1;
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&previous) != cudaSuccess) previous = -1;
    if (previous != device) cuda_check(cudaSetDevice(device), "cudaSetDevice(rt.device)", __FILE__, __LINE__);
    else previous = -1;
  }
  ~DeviceScope() {
    if (previous >= 0) cudaSetDevice(previous);
  }
};

template <class F>
inline void guarded(char const *name, F &&f) {
  try {
    std::lock_guard<std::mutex> lock(runtime().mutex);
    runtime().ensure();
    DeviceScope scope(runtime().device);
    f();
  } catch (CudaFailure const &e) {
    std::string m = std::string(name) + ": " + e.what;
    ls_hs_error(m.c_str());
  }
}

// ---- small RAII helpers -------------------------------------------------------
template <class T>
struct DeviceBuffer {
  T *ptr = nullptr;
  size_t capacity = 0;  // elements
  DeviceBuffer() = default;
  DeviceBuffer(DeviceBuffer const &) = delete;
  DeviceBuffer &operator=(DeviceBuffer const &) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    capacity = 0;
  }
  // Grow-only; contents are NOT preserved.
  T *reserve(size_t n) {
    if (n > capacity) {
      release();
      size_t want = n + n / 4 + 64;
      CUDA_CHECK(cudaMalloc(&ptr, want * sizeof(T)));
      capacity = want;
    }
    return ptr;
  }
  T *take() {
    T *p = ptr;
    ptr = nullptr;
    capacity = 0;
    return p;
  }
};

template <class T>
struct PinnedBuffer {
  T *ptr = nullptr;
  size_t capacity = 0;
  PinnedBuffer() = default;
  PinnedBuffer(PinnedBuffer const &) = delete;
  PinnedBuffer &operator=(PinnedBuffer const &) = delete;
  ~PinnedBuffer() { release(); }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    capacity = 0;
  }
  T *reserve(size_t n) {
    if (n > capacity) {
      release();
      size_t want = n + n / 4 + 64;
      CUDA_CHECK(cudaMallocHost(&ptr, want * sizeof(T)));
      capacity = want;
    }
    return ptr;
  }
};

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- device-side bit helpers ----------------------------------------------
#if defined(__CUDACC__)
// One Benes stage: swap bit pairs (j, j+d) selected by m.
template <class W>
__device__ __forceinline__ W bit_permute_step(W x, W m, unsigned d) {
  W const y = ((x >> d) ^ x) & m;
  return (x ^ y) ^ (y << d);
}
// Gosper: next integer with the same popcount.
__device__ __forceinline__ uint64_t next_same_popcount(uint64_t v) {
  uint64_t const t = v | (v - 1);
  return (t + 1) | (((~t & (t + 1)) - 1) >> (__ffsll((long long)v)));
}
#endif

}  // namespace lsb
