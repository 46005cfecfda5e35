// runtime.cu -- process runtime + the ABI glue of kernels/reference.c
// (error handler, refcounts, external arrays, the ls_chpl_kernels vtable) and
// of chapel/src/library.c (ls_chpl_init / ls_chpl_finalize).
#include <atomic>

#include <map>

#include "state.hpp"

namespace lsb {

Runtime &runtime() {
  static Runtime rt;
  return rt;
}

void Runtime::ensure() {
  if (own_stream != nullptr) return;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw CudaFailure{"no CUDA device available: liblattice_symmetries_b200 has no CPU fallback"};
  // One rank per GPU: honour the device already selected by the process
  // (torch.cuda.set_device / CUDA_VISIBLE_DEVICES), else LOCAL_RANK.
  // Was a device selected on this thread already?  cudaGetDevice() says 0 either way, so ask the driver whether a
  // context is current (cudaSetDevice binds the primary context since CUDA 12) BEFORE any runtime call creates one.
  bool selected = false;
  {
    using CtxGetCurrent = int (*)(void **);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &fn, cudaEnableDefault, &q) == cudaSuccess && fn != nullptr) {
      void *ctx = nullptr;
      if (reinterpret_cast<CtxGetCurrent>(fn)(&ctx) == 0 && ctx != nullptr) selected = true;
    }
    (void)cudaGetLastError();
  }
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (char const *s = getenv("LS_B200_DEVICE")) dev = atoi(s) % count;
  else if (char const *r = getenv("LOCAL_RANK")) {
    // torchrun, and nothing selected a device on this thread yet -> this rank's own GPU, so that the library lands on
    // the same device a later torch.cuda.set_device(LOCAL_RANK) picks.  An explicit earlier choice (device 0 included) wins.
    if (!selected && count > 1) dev = atoi(r) % count;
  }
  CUDA_CHECK(cudaSetDevice(dev));
  device = dev;
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  sm_count = prop.multiProcessorCount;
  smem_optin = prop.sharedMemPerBlockOptin;
  CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  stream = own_stream;
  CUDA_CHECK(cudaEventCreate(&ev0));
  CUDA_CHECK(cudaEventCreate(&ev1));
}

// ---- block cache ---------------------------------------------------------------------------------------------------
// Large device allocations of the library (representatives, norms, index tables, the replicated vector, build outputs)
// go through a caching layer: a freed block is kept (up to LS_B200_CACHE_GB, default 8 GB in total) and handed out
// again to the next request of about the same size.  cudaMalloc / cudaFree / cudaMallocManaged are device-wide
// synchronising driver calls whose cost on the test boxes ranges from 0.1 ms to > 1 s once NCCL has enabled peer
// access (every allocation is then mapped into all peers): a second build of the same basis, or a re-balancing, then
// never reaches the driver.  block_free accepts any device pointer (foreign ones are cudaFree'd).
namespace {
struct BlockInfo {
  size_t capacity;
  bool managed;
};
struct BlockCache {
  std::mutex mutex;
  std::unordered_map<void *, BlockInfo> live;
  std::multimap<size_t, void *> idle[2];  // [managed]
  size_t idle_bytes = 0;
  size_t limit() const {
    static size_t const value = [] {
      double gb = 8.0;
      if (char const *env = getenv("LS_B200_CACHE_GB")) gb = atof(env);
      return (size_t)(gb * (double)(size_t(1) << 30));
    }();
    return value;
  }
};
BlockCache &block_cache() {
  static BlockCache *c = new BlockCache();  // never destroyed: finalizers may run after static destruction began
  return *c;
}
constexpr size_t kCacheMinBytes = size_t(256) << 10;
}  // namespace

size_t block_cache_trim(size_t keep_bytes) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mutex);
  size_t released = 0;
  for (int m = 0; m < 2; ++m) {
    while (!c.idle[m].empty() && c.idle_bytes > keep_bytes) {
      auto it = std::prev(c.idle[m].end());  // largest first
      cudaFree(it->second);
      released += it->first;
      c.idle_bytes -= it->first;
      c.idle[m].erase(it);
    }
  }
  return released;
}

size_t block_cache_idle_bytes() {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mutex);
  return c.idle_bytes;
}

void *block_alloc(size_t bytes, bool managed) {
  Runtime &rt = runtime();
  BlockCache &c = block_cache();
  bytes = std::max<size_t>(bytes, 8);
  if (bytes >= kCacheMinBytes) {
    std::lock_guard<std::mutex> lock(c.mutex);
    auto &idle = c.idle[managed ? 1 : 0];
    auto it = idle.lower_bound(bytes);
    if (it != idle.end() && it->first <= bytes + bytes / 4 + (size_t(2) << 20)) {
      void *p = it->second;
      size_t const cap = it->first;
      idle.erase(it);
      c.idle_bytes -= cap;
      c.live[p] = BlockInfo{cap, managed};
      if (managed) {
        // the previous owner advised read-mostly once its contents were final: the next one writes from the device
        (void)cudaMemAdvise(p, cap, cudaMemAdviseUnsetReadMostly, rt.device);
        (void)cudaMemPrefetchAsync(p, cap, rt.device, rt.stream);
        (void)cudaGetLastError();
      }
      return p;
    }
  }
  size_t const cap = bytes >= kCacheMinBytes ? (bytes + (size_t(2) << 20) - 1) / (size_t(2) << 20) * (size_t(2) << 20) : bytes;
  void *p = nullptr;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cudaError_t const e = managed ? cudaMallocManaged(&p, cap) : cudaMalloc(&p, cap);
    if (e == cudaSuccess) break;
    (void)cudaGetLastError();
    p = nullptr;
    if (attempt == 0 && block_cache_trim(0) > 0) continue;  // give the idle blocks back and try once more
    if (managed) return nullptr;                            // callers fall back to plain device memory
    cuda_check(e, "cudaMalloc (block_alloc)", __FILE__, __LINE__);
  }
  if (managed) {
    CUDA_CHECK(cudaMemAdvise(p, cap, cudaMemAdviseSetPreferredLocation, rt.device));
    CUDA_CHECK(cudaMemPrefetchAsync(p, cap, rt.device, rt.stream));
  }
  if (cap >= kCacheMinBytes) {
    std::lock_guard<std::mutex> lock(c.mutex);
    c.live[p] = BlockInfo{cap, managed};
  }
  return p;
}

void block_free(void *p) {
  if (p == nullptr) return;
  BlockCache &c = block_cache();
  {
    std::lock_guard<std::mutex> lock(c.mutex);
    auto it = c.live.find(p);
    if (it != c.live.end()) {
      BlockInfo const info = it->second;
      c.live.erase(it);
      if (info.capacity <= c.limit() && c.idle_bytes + info.capacity <= c.limit()) {
        c.idle[info.managed ? 1 : 0].emplace(info.capacity, p);
        c.idle_bytes += info.capacity;
        return;
      }
    }
  }
  cudaFree(p);
}

// Allocation of the large randomly-accessed tables (index levels, keys, the replicated vector): the block cache, or --
// LS_B200_POOL_MALLOC=1, an A/B knob -- stream-ordered allocations from the device's default memory pool, which unlike
// cudaMalloc under peer access is mapped by this device only.  A/B on 2 x B200 (chain-40, 17 GB of tables, random 8-byte
// gathers): no difference in kernel time (397 ms either way).
void *alloc_local(size_t bytes) {
  static bool const pool = getenv("LS_B200_POOL_MALLOC") != nullptr;
  if (!pool) return block_alloc(bytes, false);
  Runtime &rt = runtime();
  void *p = nullptr;
  CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(bytes, 8), rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  return p;
}

std::unordered_map<void const *, BuiltReps> &built_registry() {
  static std::unordered_map<void const *, BuiltReps> m;
  return m;
}

}  // namespace lsb

using namespace lsb;

extern "C" {

// ---- kernels/reference.c:11-36 ---------------------------------------------
void ls_hs_fatal_error(char const *func, int const line, char const *message) {
  fprintf(stderr, "[Error]   [%s#%i] %s\n[Error]   Aborting ...", func, line, message);
  abort();
}

typedef void (*error_handler_type)(char const *);
static void default_error_handler(char const *message) {
  fprintf(stderr, "[Error]   %s\n[Error]   Aborting ...", message);
  abort();
}
static std::atomic<error_handler_type> g_error_handler{default_error_handler};

void ls_hs_set_exception_handler(error_handler_type handler) {
  if (handler == nullptr) handler = default_error_handler;
  g_error_handler.store(handler);
}

void ls_hs_error(char const *message) {
  error_handler_type handler = g_error_handler.load();
  LSB_CHECK(handler != nullptr, "error handler is NULL");
  (*handler)(message);
}

// ---- kernels/reference.c:40-64 ---------------------------------------------
void ls_hs_internal_destroy_external_array(chpl_external_array *arr) {
  LSB_CHECK(arr != nullptr, "trying to destroy a NULL chpl_external_array");
  if (arr->freer != nullptr) {
    auto const free_func = reinterpret_cast<void (*)(void *)>(arr->freer);
    (*free_func)(arr->elts);
  }
}

int ls_hs_internal_read_refcount(int const *refcount) {
  return __atomic_load_n(refcount, __ATOMIC_SEQ_CST);
}
void ls_hs_internal_write_refcount(int *refcount, int value) {
  __atomic_store_n(refcount, value, __ATOMIC_SEQ_CST);
}
int ls_hs_internal_inc_refcount(int *refcount) {
  return __atomic_fetch_add(refcount, 1, __ATOMIC_SEQ_CST);
}
int ls_hs_internal_dec_refcount(int *refcount) {
  return __atomic_fetch_sub(refcount, 1, __ATOMIC_SEQ_CST);
}

// ---- kernels/reference.c:214-225 -------------------------------------------
static ls_chpl_kernels g_chpl_kernels = {nullptr, nullptr, nullptr, nullptr};
ls_chpl_kernels const *ls_hs_internal_get_chpl_kernels(void) { return &g_chpl_kernels; }
void ls_hs_internal_set_chpl_kernels(ls_chpl_kernels const *kernels) { g_chpl_kernels = *kernels; }

// ---- chapel/src/LatticeSymmetries.chpl:18-33, chapel/src/library.c:19-34 ----
void ls_chpl_init_kernels(void) {
  ls_chpl_kernels k;
  k.enumerate_states = &ls_chpl_enumerate_representatives;
  k.operator_apply_off_diag = &ls_chpl_operator_apply_off_diag;
  k.operator_apply_diag = &ls_chpl_operator_apply_diag;
  k.matrix_vector_product = &ls_chpl_matrix_vector_product;
  ls_hs_internal_set_chpl_kernels(&k);
}

void ls_chpl_init(void) {
  // The Chapel runtime boots here in the reference; we register the vtable and
  // bring up the device eagerly so that a missing GPU is reported at init.
  ls_chpl_init_kernels();
  guarded("ls_chpl_init", [] {});
}

void ls_chpl_finalize(void) {
  // The reference's chpl_library_finalize() ends in exit(0)
  // (python/lattice_symmetries/__init__.py:58); we just drain the stream.
  Runtime &rt = runtime();
  if (rt.stream != nullptr) cudaStreamSynchronize(rt.stream);
}

// ---- extensions ---------------------------------------------------------------
uint64_t ls_b200_kernel_launch_count(void) { return runtime().launches.load(); }

double ls_b200_last_kernel_ms(char const *name) {
  Runtime &rt = runtime();
  if (name != nullptr && strcmp(name, "build") == 0) return rt.last_build_ms;
  if (name != nullptr && strcmp(name, "orbit") == 0) return rt.last_orbit_ms;
  if (name != nullptr && strcmp(name, "gather") == 0) return rt.last_gather_ms;
  if (name != nullptr && strcmp(name, "combine") == 0) return rt.last_combine_ms;
  if (name != nullptr && strcmp(name, "count") == 0) return rt.last_count_ms;  // row_count + scan
  if (name != nullptr && strcmp(name, "allgather") == 0) return rt.last_allgather_ms;  // of the product BEFORE the last one
  if (name != nullptr && strcmp(name, "orbit_launches") == 0) return (double)rt.last_orbit_launches;
  if (name != nullptr && strcmp(name, "gather_launches") == 0) return (double)rt.last_gather_launches;
  return rt.last_matvec_ms;
}

void *ls_b200_stream(void) {
  void *s = nullptr;
  guarded("ls_b200_stream", [&] { s = (void *)runtime().stream; });
  return s;
}

int ls_b200_set_stream(void *cuda_stream) {
  int status = -1;
  guarded("ls_b200_set_stream", [&] {
    Runtime &rt = runtime();
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    rt.stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : rt.own_stream;
    status = 0;
  });
  return status;
}

int ls_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void *ls_b200_device_malloc(size_t bytes) {
  void *p = nullptr;
  guarded("ls_b200_device_malloc", [&] { p = block_alloc(bytes, false); });
  return p;
}
void ls_b200_device_free(void *p) {
  if (p != nullptr) block_free(p);
}

int ls_b200_copy_to_device(void *dst_dev, void const *src_host, size_t bytes) {
  int status = -1;
  guarded("ls_b200_copy_to_device", [&] {
    if (bytes > 0) {
      CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, runtime().stream));
      CUDA_CHECK(cudaStreamSynchronize(runtime().stream));
    }
    status = 0;
  });
  return status;
}
int ls_b200_copy_to_host(void *dst_host, void const *src_dev, size_t bytes) {
  int status = -1;
  guarded("ls_b200_copy_to_host", [&] {
    if (bytes > 0) {
      CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, runtime().stream));
      CUDA_CHECK(cudaStreamSynchronize(runtime().stream));
    }
    status = 0;
  });
  return status;
}
void *ls_b200_host_malloc(size_t bytes) {
  void *p = nullptr;
  guarded("ls_b200_host_malloc", [&] { CUDA_CHECK(cudaMallocHost(&p, bytes)); });
  return p;
}
void ls_b200_host_free(void *p) {
  if (p != nullptr) cudaFreeHost(p);
}

}  // extern "C"
