// index.cu -- state -> index ranking over the sorted representatives.
// Replaces kernels/indexing.c: the prefix-bucket table (:45-117) is built on
// the device by one lower_bound per prefix, and the search kernel (:273-325)
// runs one needle per thread.  Results (index or -1) are identical.
#include <cub/device/device_scan.cuh>

#include "state.hpp"

namespace lsb {

IndexView IndexData::view() const {
  IndexView v{};
  v.reps = d_reps;
  v.number_states = number_states;
  v.offsets32 = d_offsets32;
  v.offsets64 = d_offsets64;
  v.lows16 = d_lows16;
  v.lows32 = d_lows32;
  v.low_mask = shift >= 64 ? ~uint64_t(0) : ((uint64_t(1) << shift) - 1);
  v.shift = shift;
  v.identity = identity ? 1 : 0;
  v.number_buckets = uint64_t(1) << prefix_bits;
  v.steps = steps;
  v.sub_info = d_sub_info;
  v.subtab = d_subtab;
  v.entry8 = d_entry8;
  return v;
}

IndexData::~IndexData() {
  matvec_forget_index(this);
  delete dist;
  if (owns_d_reps) block_free(d_reps);
  block_free(d_offsets32);
  block_free(d_offsets64);
  block_free(d_lows16);
  block_free(d_lows32);
  block_free(d_sub_info);
  block_free(d_subtab);
  block_free(d_entry8);
  block_free(d_norms);
  magic = 0;
}

// offsets[p] = first index whose (rep >> shift) >= p  (kernels/indexing.c:63-69)
template <class T>
__global__ void __launch_bounds__(256)
bucket_offsets_kernel(uint64_t const *__restrict__ reps, int64_t n, int shift, int64_t number_offsets,
                      T *__restrict__ offsets) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < number_offsets;
       p += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n;
    if (p == number_offsets - 1) {
      lo = n;
    } else {
      uint64_t const key = (uint64_t)p;
      while (lo < hi) {
        int64_t const mid = (lo + hi) >> 1;
        if ((__ldg(reps + mid) >> shift) < key) lo = mid + 1; else hi = mid;
      }
    }
    offsets[p] = (T)lo;
  }
}

template <class T>
__global__ void __launch_bounds__(256)
low_bits_kernel(uint64_t const *__restrict__ reps, int64_t n, uint64_t mask, T *__restrict__ lows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    lows[i] = (T)(reps[i] & mask);
}

// Largest bucket of the table (sets the trip count of the branchless search).
template <class T>
__global__ void __launch_bounds__(256)
max_bucket_kernel(T const *__restrict__ offsets, int64_t number_buckets, unsigned long long *__restrict__ out) {
  unsigned long long local = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < number_buckets;
       p += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long const w = (unsigned long long)(offsets[p + 1] - offsets[p]);
    local = w > local ? w : local;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long const other = __shfl_xor_sync(0xffffffffu, local, o);
    local = other > local ? other : local;
  }
  if ((threadIdx.x & 31) == 0 && local != 0) atomicMax(out, local);
}

// ---- second level for crowded buckets ---------------------------------------------
#ifndef LS_INDEX_DENSE
#define LS_INDEX_DENSE 15
#endif
#ifndef LS_INDEX_SUBLOG2
#define LS_INDEX_SUBLOG2 3
#endif
constexpr int kDenseBucket = LS_INDEX_DENSE;  // buckets with more entries get a sub-table ...
constexpr int kSubTargetLog2 = LS_INDEX_SUBLOG2; // ... of about 2^3 entries per slot
constexpr int kMaxSubBits = 24;

__device__ __forceinline__ int sub_bits(uint32_t n, int shift) {
  // (positions inside a sub-table are packed into 26 bits; larger buckets -- never seen -- are searched directly)
  if (n <= (uint32_t)kDenseBucket || shift <= 0 || n >= (1u << 26)) return 0;
  int const len = 32 - __clz(n - 1);  // ceil(log2 n)
  return max(1, min(min(shift, kMaxSubBits), len - kSubTargetLog2));
}

// units[p] = size of bucket p's sub-table in units of 8 entries (0: none)
__global__ void __launch_bounds__(256)
sub_sizes_kernel(uint32_t const *__restrict__ offsets, int64_t number_buckets, int shift,
                 uint32_t *__restrict__ units) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < number_buckets;
       p += (int64_t)gridDim.x * blockDim.x) {
    int const p2 = sub_bits(offsets[p + 1] - offsets[p], shift);
    units[p] = p2 == 0 ? 0u : (((1u << p2) + 1u + 7u) >> 3);
  }
}

// Warp per bucket: sub_info[p] and the lower bounds of its 2^p2 + 1 slot boundaries;
// also the largest window any lookup can end up with.  `units` is overwritten in
// place by sub_info (first[p] is read before).
__global__ void __launch_bounds__(256)
sub_fill_kernel(IndexView ix, uint32_t const *__restrict__ first, uint32_t *__restrict__ units_then_info,
                uint32_t *__restrict__ subtab, uint2 *__restrict__ entry8, unsigned long long *__restrict__ max_window) {
  int const lane = threadIdx.x & 31;
  int64_t const warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t const warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  bool const compact = ix.lows16 != nullptr || ix.lows32 != nullptr;
  uint32_t widest = 0;
  for (int64_t p = warp; p < (int64_t)ix.number_buckets; p += warps) {
    uint32_t const lo = ix.offsets32[p];
    uint32_t const n = ix.offsets32[p + 1] - lo;
    int const p2 = sub_bits(n, ix.shift);
    if (p2 == 0) {
      widest = max(widest, n);
      if (lane == 0) {
        units_then_info[p] = 0;
        entry8[p] = make_uint2(lo, n);  // a plain length: n <= kDenseBucket, or an oversized bucket below 2^27
      }
      continue;
    }
    uint32_t const t = first[p];
    uint32_t *tab = subtab + (size_t)t * 8;
    uint32_t const slots = 1u << p2;
    uint64_t const high = compact ? 0 : ((uint64_t)p << ix.shift);
    for (uint32_t e = lane; e <= slots; e += 32) {
      uint32_t pos = n;
      if (e < slots) {
        uint64_t const key = high | ((uint64_t)e << (ix.shift - p2));
        uint32_t a = 0, b = n;
        while (a < b) {
          uint32_t const mid = (a + b) >> 1;
          if (index_key_at(ix, (int64_t)lo + mid) < key) a = mid + 1; else b = mid;
        }
        pos = a;
      }
      tab[e] = pos;
    }
    __syncwarp();
    for (uint32_t e = lane; e < slots; e += 32) widest = max(widest, tab[e + 1] - tab[e]);
    __syncwarp();
    // pack: position << 6 | min(length, 63), 32 entries at a time in ascending order (an entry's length needs its
    // right neighbour still unpacked)
    for (uint32_t base = 0; base <= slots; base += 32) {
      uint32_t const e = base + lane;
      uint32_t v0 = 0, v1 = 0;
      if (e <= slots) {
        v0 = tab[e];
        v1 = e < slots ? tab[e + 1] : v0;
      }
      __syncwarp();
      if (e <= slots) tab[e] = (v0 << 6) | min(v1 - v0, 63u);
      __syncwarp();
    }
    if (lane == 0) {
      units_then_info[p] = ((uint32_t)p2 << 27) | t;
      entry8[p] = make_uint2(lo, ((uint32_t)p2 << 27) | t);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) widest = max(widest, __shfl_xor_sync(0xffffffffu, widest, o));
  if (lane == 0 && widest != 0) atomicMax(max_window, (unsigned long long)widest);
}

__global__ void __launch_bounds__(256)
state_index_kernel(IndexView ix, int64_t n, uint64_t const *__restrict__ needles,
                   int64_t *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = state_index(ix, needles[i]);
}

// Adaptive prefix width: about eight representatives per bucket (the bucket
// table then is ~1/16 of the low-bits array and both stay L2-resident), widened
// when that lets the low bits fit 16-bit keys.  The reference's caller-chosen
// width (22 or 26, kernels/reference.c:188, chapel/src/CommonParameters.chpl:7)
// only affects speed, never results, so we are free to choose.
static int choose_prefix_bits(int64_t n, int number_bits, int requested) {
  (void)requested;
  if (char const *env = getenv("LS_B200_INDEX_PREFIX")) return std::max(0, std::min(atoi(env), number_bits));  // A/B knob
  int want = 1;
  while (want < 40 && (int64_t(1) << want) < n / 8) ++want;
  int const cap = 28;
  int p = std::min(want, cap);
  if (number_bits - p > 16 && number_bits - 16 <= std::min(cap, want + 3)) p = number_bits - 16;
  p = std::min(p, number_bits);
  return std::max(p, 0);
}

// LS_B200_DIST_LEAN_LOCAL=1: the LOCAL index of a sharded basis is a 256-bucket table over the representatives
// themselves -- no compact keys, no second level (~6 bytes per state saved).  All-gather products never search it; host
// queries (ls_hs_state_index) and all-to-all products still work, with longer searches.
static bool g_lean_index = false;
void index_set_lean(bool lean) { g_lean_index = lean; }

void build_bucket_table(IndexData &ix, int requested_prefix_bits) {
  Runtime &rt = runtime();
  ix.prefix_bits = choose_prefix_bits(ix.number_states, ix.number_bits, requested_prefix_bits);
  bool const lean = g_lean_index;
  if (lean) ix.prefix_bits = std::min(8, ix.number_bits);
  ix.shift = ix.number_bits - ix.prefix_bits;
  int64_t const number_offsets = (int64_t(1) << ix.prefix_bits) + 1;
  unsigned const blocks = (unsigned)std::min<int64_t>((number_offsets + 255) / 256, (int64_t)rt.sm_count * 16);
  if (ix.number_states < (int64_t(1) << 32)) {
    alloc_local(&ix.d_offsets32, sizeof(uint32_t) * (size_t)number_offsets);
    bucket_offsets_kernel<uint32_t><<<blocks, 256, 0, rt.stream>>>(ix.d_reps, ix.number_states, ix.shift,
                                                                 number_offsets, ix.d_offsets32);
  } else {
    alloc_local(&ix.d_offsets64, sizeof(int64_t) * (size_t)number_offsets);
    bucket_offsets_kernel<int64_t><<<blocks, 256, 0, rt.stream>>>(ix.d_reps, ix.number_states, ix.shift,
                                                                number_offsets, ix.d_offsets64);
  }
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  uint64_t const mask = ix.shift >= 64 ? ~uint64_t(0) : ((uint64_t(1) << ix.shift) - 1);
  unsigned const lblocks = (unsigned)std::min<int64_t>((ix.number_states + 255) / 256, (int64_t)rt.sm_count * 16);
  if (lean) {
    // (shift > 32: the searches compare whole representatives)
  } else if (ix.shift <= 16) {
    alloc_local(&ix.d_lows16, sizeof(uint16_t) * (size_t)ix.number_states);
    low_bits_kernel<uint16_t><<<lblocks, 256, 0, rt.stream>>>(ix.d_reps, ix.number_states, mask, ix.d_lows16);
    count_launch();
  } else if (ix.shift <= 32) {
    alloc_local(&ix.d_lows32, sizeof(uint32_t) * (size_t)ix.number_states);
    low_bits_kernel<uint32_t><<<lblocks, 256, 0, rt.stream>>>(ix.d_reps, ix.number_states, mask, ix.d_lows32);
    count_launch();
  }
  // temporaries live for the process (grow-only): a cudaFree inside the build is a device-wide synchronisation
  // that cost up to 200 ms on the test boxes
  static DeviceBuffer<unsigned long long> max_buffer;
  static DeviceBuffer<uint32_t> first_buffer;
  static DeviceBuffer<unsigned char> scan_buffer;
  unsigned long long *d_max = max_buffer.reserve(1), h_max = 0;
  CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof h_max, rt.stream));
  static bool const no_sub = getenv("LS_B200_INDEX_FLAT") != nullptr;  // A/B knob: first level only
  bool have_sub = false;
  if (ix.d_offsets32 != nullptr && ix.shift > 0 && !no_sub && !lean) {
    // second level: sizes -> exclusive scan -> fill; dropped when nothing is crowded
    int64_t const nb = number_offsets - 1;
    uint32_t *d_units = nullptr;
    alloc_local(&d_units, sizeof(uint32_t) * (size_t)(nb + 1));  // becomes sub_info
    uint32_t *d_first = first_buffer.reserve((size_t)(nb + 1));
    CUDA_CHECK(cudaMemsetAsync(d_units + nb, 0, sizeof(uint32_t), rt.stream));
    sub_sizes_kernel<<<blocks, 256, 0, rt.stream>>>(ix.d_offsets32, nb, ix.shift, d_units);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_units, d_first, (int)(nb + 1), rt.stream);
    void *d_tmp = scan_buffer.reserve(tmp_bytes);
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_units, d_first, (int)(nb + 1), rt.stream);
    count_launch(2);
    uint32_t total_units = 0;
    CUDA_CHECK(cudaMemcpyAsync(&total_units, d_first + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    // (no 32-bit wrap: a crowded bucket of n states adds at most n / 64 + 1 units and n sums to < 2^32)
    if (total_units > 0 && total_units < (1u << 27)) {
      alloc_local(&ix.d_subtab, sizeof(uint32_t) * (size_t)total_units * 8);
      alloc_local(&ix.d_entry8, sizeof(uint2) * (size_t)nb);
      IndexView v = ix.view();
      v.sub_info = nullptr;
      v.entry8 = nullptr;
      sub_fill_kernel<<<blocks, 256, 0, rt.stream>>>(v, d_first, d_units, ix.d_subtab, ix.d_entry8, d_max);
      count_launch();
      ix.d_sub_info = d_units;
      d_units = nullptr;
      have_sub = true;
    }
    if (d_units != nullptr) block_free(d_units);  // nothing was crowded
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  }
  if (have_sub) {
    // sub_fill_kernel has folded every final window into d_max
  } else if (ix.d_offsets32 != nullptr) {
    max_bucket_kernel<uint32_t><<<blocks, 256, 0, rt.stream>>>(ix.d_offsets32, number_offsets - 1, d_max);
    count_launch();
  } else {
    max_bucket_kernel<int64_t><<<blocks, 256, 0, rt.stream>>>(ix.d_offsets64, number_offsets - 1, d_max);
    count_launch();
  }
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof h_max, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  ix.steps = 0;
  while ((h_max >> ix.steps) != 0) ++ix.steps;
}

IndexData *index_of(ls_hs_basis const *basis) {
  if (basis == nullptr || basis->kernels == nullptr) return nullptr;
  auto *ix = static_cast<IndexData *>(basis->kernels->state_index_data);
  if (ix == nullptr || ix->magic != kIndexMagic) return nullptr;
  return ix;
}

// Creates the index object from host representatives (or adopts the device
// copy a just-finished build left in the registry).
IndexData *create_index(uint64_t const *host_reps, int64_t count, int number_bits, int prefix_bits) {
  auto *ix = new IndexData();
  ix->number_states = count;
  ix->host_reps = host_reps;
  ix->number_bits = number_bits;
  auto &reg = built_registry();
  auto it = host_reps ? reg.find(host_reps) : reg.end();
  if (it != reg.end() && (int64_t)it->second.count == count) {
    ix->d_reps = it->second.d_reps;
    ix->d_norms = it->second.d_norms;
    // a managed array doubles as the caller's host view and is released by its freer
    ix->owns_d_reps = static_cast<void const *>(ix->d_reps) != static_cast<void const *>(host_reps);
    reg.erase(it);
  } else if (count > 0) {
    block_alloc(&ix->d_reps, sizeof(uint64_t) * (size_t)count);
    CUDA_CHECK(cudaMemcpyAsync(ix->d_reps, host_reps, sizeof(uint64_t) * (size_t)count,
                               cudaMemcpyHostToDevice, runtime().stream));
  }
  if (count > 0 && number_bits > 0) {
    build_bucket_table(*ix, prefix_bits);
  } else {
    while ((ix->number_states >> ix->steps) != 0) ++ix->steps;
  }
  CUDA_CHECK(cudaStreamSynchronize(runtime().stream));
  return ix;
}

// ---- pieces of the table build used by the replicated index of a sharded basis (dist.cu) ----------------------
int index_choose_prefix_bits(int64_t n, int number_bits) { return choose_prefix_bits(n, number_bits, 0); }

// keys[i] = low `shift` bits of reps[i], as 2- or 4-byte keys
void index_local_keys(uint64_t const *d_reps, int64_t n, int shift, void *d_keys, int key_bytes) {
  if (n <= 0) return;
  Runtime &rt = runtime();
  uint64_t const mask = shift >= 64 ? ~uint64_t(0) : ((uint64_t(1) << shift) - 1);
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 16);
  if (key_bytes == 2) low_bits_kernel<uint16_t><<<blocks, 256, 0, rt.stream>>>(d_reps, n, mask, static_cast<uint16_t *>(d_keys));
  else low_bits_kernel<uint32_t><<<blocks, 256, 0, rt.stream>>>(d_reps, n, mask, static_cast<uint32_t *>(d_keys));
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

// out[p] = number of LOCAL representatives whose prefix is below p (p = 0 .. number_offsets - 1; the last entry is
// n): summed over the ranks this is the level-1 table of the whole basis.
void index_local_offsets64(uint64_t const *d_reps, int64_t n, int shift, int64_t number_offsets, int64_t *d_out) {
  Runtime &rt = runtime();
  unsigned const blocks = (unsigned)std::min<int64_t>((number_offsets + 255) / 256, (int64_t)rt.sm_count * 16);
  bucket_offsets_kernel<int64_t><<<blocks, 256, 0, rt.stream>>>(d_reps, n, shift, number_offsets, d_out);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

int index_steps_from_offsets64(int64_t const *d_offsets, int64_t number_buckets) {
  Runtime &rt = runtime();
  static DeviceBuffer<unsigned long long> max_buffer;
  unsigned long long *d_max = max_buffer.reserve(1), h_max = 0;
  CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof h_max, rt.stream));
  unsigned const blocks = (unsigned)std::min<int64_t>((number_buckets + 255) / 256, (int64_t)rt.sm_count * 16);
  max_bucket_kernel<int64_t><<<blocks, 256, 0, rt.stream>>>(d_offsets, number_buckets, d_max);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof h_max, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  int steps = 0;
  while ((h_max >> steps) != 0) ++steps;
  return steps;
}

// Index over a device-resident sorted list the object takes ownership of (no host view).
IndexData *create_index_from_device(uint64_t *d_reps, int64_t count, int number_bits, int prefix_bits) {
  auto *ix = new IndexData();
  ix->number_states = count;
  ix->host_reps = nullptr;
  ix->number_bits = number_bits;
  ix->d_reps = d_reps;
  ix->owns_d_reps = true;
  if (count > 0 && number_bits > 0) {
    build_bucket_table(*ix, prefix_bits);
  } else {
    while ((ix->number_states >> ix->steps) != 0) ++ix->steps;
  }
  CUDA_CHECK(cudaStreamSynchronize(runtime().stream));
  return ix;
}

// out[i] = position of needles[i] in the index (device pointers), -1 when absent; on the library stream
void launch_state_index(IndexData const &ix, int64_t n, uint64_t const *d_needles, int64_t *d_out) {
  if (n <= 0) return;
  Runtime &rt = runtime();
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 16);
  state_index_kernel<<<blocks, 256, 0, rt.stream>>>(ix.view(), n, d_needles, d_out);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

struct IndexScratch {
  DeviceBuffer<uint64_t> needles;
  DeviceBuffer<int64_t> out;
};
static IndexScratch &index_scratch() {
  static IndexScratch s;
  return s;
}

}  // namespace lsb

using namespace lsb;

extern "C" {

ls_hs_state_index_binary_search_data *ls_hs_create_state_index_binary_search_kernel_data(
    chpl_external_array const *representatives, int const number_bits, int const prefix_bits) {
  IndexData *ix = nullptr;
  guarded(__func__, [&] {
    ix = create_index(static_cast<uint64_t const *>(representatives->elts),
                      (int64_t)representatives->num_elts, number_bits, prefix_bits);
  });
  return reinterpret_cast<ls_hs_state_index_binary_search_data *>(ix);
}

void ls_hs_destroy_state_index_binary_search_kernel_data(ls_hs_state_index_binary_search_data *cache) {
  if (cache == nullptr) return;
  auto *ix = reinterpret_cast<IndexData *>(cache);
  LSB_CHECK(ix->magic == kIndexMagic, "not an index object of this library");
  std::lock_guard<std::mutex> lock(runtime().mutex);
  delete ix;
}

int ls_b200_index_info(ls_hs_basis const *basis, int64_t out[4]) {
  IndexData const *ix = index_of(basis);
  if (ix == nullptr || out == nullptr) return -1;
  out[0] = ix->prefix_bits;
  out[1] = ix->steps;
  out[2] = ix->d_sub_info != nullptr ? 1 : 0;
  out[3] = ix->number_states;
  return 0;
}

void ls_hs_state_index_binary_search_kernel(ptrdiff_t const batch_size, uint64_t const *spins,
                                            ptrdiff_t const spins_stride, ptrdiff_t *indices,
                                            ptrdiff_t const indices_stride,
                                            void const *private_kernel_data) {
  auto const *ix = static_cast<IndexData const *>(private_kernel_data);
  LSB_CHECK(ix != nullptr && ix->magic == kIndexMagic, "invalid state_index kernel data");
  LSB_CHECK(indices_stride == 1, "expected indices_stride==1");  // indexing.c:280
  LSB_CHECK(spins_stride == 1, "expected spins_stride==1");      // indexing.c:281
  static_assert(sizeof(ptrdiff_t) == sizeof(int64_t), "ptrdiff_t must be 64-bit");
  guarded(__func__, [&] {
    IndexScratch &sc = index_scratch();
    Runtime &rt = runtime();
    constexpr int64_t chunk = int64_t(1) << 22;
    for (int64_t begin = 0; begin < batch_size; begin += chunk) {
      int64_t const n = std::min<int64_t>(chunk, batch_size - begin);
      uint64_t *d_in = sc.needles.reserve((size_t)n);
      int64_t *d_out = sc.out.reserve((size_t)n);
      CUDA_CHECK(cudaMemcpyAsync(d_in, spins + begin, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, rt.stream));
      unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 16);
      state_index_kernel<<<blocks, 256, 0, rt.stream>>>(ix->view(), n, d_in, d_out);
      count_launch();
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaMemcpyAsync(indices + begin, d_out, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, rt.stream));
      CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    }
  });
}

}  // extern "C"
