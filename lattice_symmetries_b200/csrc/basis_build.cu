// basis_build.cu -- symmetry-adapted basis construction on the device.
//
// Replaces the Chapel enumeration driver chapel/src/StatesEnumeration.chpl
// (Gosper stepping :29-32, range split :130-189, projected loop :242-268,
// unprojected :269-289, spinful product :290-326, export :692-709) and the
// Haskell combinadics it calls (haskell/src/LatticeSymmetries/Basis.hs:487-550,
// bounds :734-774).
//
// Design (not a port):
//   * candidates are addressed by their combinadic (or linear) index, so any
//     contiguous index range is an independent unit of work -- the same split
//     serves thread blocks and GPUs (ranks take contiguous index ranges and
//     their outputs concatenate in rank order, globally sorted);
//   * pass A ("flags"): every thread unranks its first candidate, Gosper-steps
//     32 of them, transposes them into bit planes (bitslice.cuh) and tests all
//     32 against the whole group with two LOP3 per plane per element; the
//     result is a 32-bit survivor mask per thread (1 bit per candidate in HBM,
//     candidates themselves are never materialised);
//   * a device-wide exclusive scan of per-block survivor counts;
//   * pass C ("scatter"): survivors are regenerated and written, in ascending
//     order, together with their norms sqrt(n/|G|).
#include <cub/device/device_scan.cuh>

#include <chrono>

#include <algorithm>

#include "bitslice.cuh"
#include "plane_table.cuh"
#include "state.hpp"

namespace lsb {

// ---- host combinadics (Basis.hs:487-550) ---------------------------------------
static uint64_t g_binom[65][65];
static bool g_binom_ready = false;
static void init_binomials() {
  if (g_binom_ready) return;
  memset(g_binom, 0, sizeof g_binom);
  for (int n = 0; n <= 64; ++n) {
    g_binom[n][0] = 1;
    for (int k = 1; k <= n; ++k) g_binom[n][k] = g_binom[n - 1][k - 1] + (k <= n - 1 ? g_binom[n - 1][k] : 0);
  }
  g_binom_ready = true;
}
uint64_t binomial(int n, int k) {
  init_binomials();
  if (n < 0 || k < 0 || k > n) return 0;
  return g_binom[n][k];
}
uint64_t fixed_hamming_state_to_index(uint64_t state) {
  init_binomials();
  uint64_t idx = 0;
  int k = 1;
  while (state != 0) {
    int const c = __builtin_ctzll(state);
    state &= state - 1;
    idx += binomial(c, k);
    ++k;
  }
  return idx;
}
uint64_t fixed_hamming_index_to_state(uint64_t index, int hamming_weight) {
  init_binomials();
  uint64_t state = 0;
  int c = 64;
  for (int i = hamming_weight; i > 0; --i) {
    do { --c; } while (binomial(c, i) > index);
    state |= uint64_t(1) << c;
    index -= binomial(c, i);
  }
  return state;
}

static uint64_t low_ones(int h) { return h >= 64 ? ~uint64_t(0) : ((uint64_t(1) << h) - 1); }

// Basis.hs:701-774 predicates and bounds, derived from the C struct alone.
BasisInfo basis_info(ls_hs_basis const *b) {
  BasisInfo info{};
  int const n = b->number_sites;
  info.number_bits = (b->particle_type == LS_HS_SPINFUL_FERMION ? 2 : 1) * n;
  GroupData const *g = nullptr;
  if (b->kernels != nullptr) {
    g = static_cast<GroupData const *>(b->kernels->is_representative_data);
    if (g != nullptr && g->magic != kGroupMagic) g = nullptr;
  }
  info.group = g;
  info.has_permutation_symmetries = b->particle_type == LS_HS_SPIN && g != nullptr && g->number_masks > 0;
  info.has_spin_inversion = b->particle_type == LS_HS_SPIN && b->spin_inversion != 0;
  info.spinful_sectors = b->particle_type == LS_HS_SPINFUL_FERMION && b->number_up != -1;
  if (b->particle_type == LS_HS_SPIN) {
    info.fixed_hamming = b->number_up != -1;
    info.hamming_weight = b->number_up;
  } else {
    info.fixed_hamming = b->number_particles != -1;
    info.hamming_weight = b->number_particles;
  }
  if (info.spinful_sectors) {
    int const up = b->number_up, down = b->number_particles - b->number_up;
    info.min_state = (low_ones(down) << n) | low_ones(up);
    info.max_state = ((low_ones(down) << (n - down)) << n) | (low_ones(up) << (n - up));
  } else if (info.fixed_hamming) {
    int const h = info.hamming_weight;
    info.min_state = low_ones(h);
    if (info.has_spin_inversion)
      info.max_state = h == 0 ? 0 : low_ones(h) << (info.number_bits - h - 1);
    else
      info.max_state = low_ones(h) << (info.number_bits - h);
  } else {
    info.min_state = 0;
    info.max_state = low_ones(info.number_bits);
  }
  return info;
}

// ---- enumeration range ------------------------------------------------------------
// Candidate k in [0, total) is
//   fixed Hamming weight : unrank(rank_min + k, hw)
//   otherwise            : min_state + k
//   spinful (up, down)   : (unrank(k / count_a, down) << n) | unrank(k % count_a, up)
struct EnumView {
  int mode;  // 0 linear, 1 fixed hamming, 2 spinful sectors
  int number_bits;
  int hw;       // hamming weight (mode 1) / n_up (mode 2)
  int hw_b;     // n_down (mode 2)
  int n_sites;  // mode 2
  uint64_t rank_min;
  uint64_t min_state;
  uint64_t count_a;  // mode 2
  uint64_t total;
  uint64_t const *binom;  // device [65][65]
  // Block-cyclic share of a rank (dist.cu): the kernels walk VIRTUAL words 0, 1, 2, ...; virtual block v >> cyc_shift
  // is block (v >> cyc_shift) * cyc_world + cyc_rank of the candidate range.  cyc_world == 0: virtual == actual.
  int cyc_shift;
  uint32_t cyc_world, cyc_rank;
};

// word of 32 candidates addressed by the kernels -> word of the candidate range
__host__ __device__ __forceinline__ uint64_t actual_word(EnumView const &e, uint64_t v) {
  if (e.cyc_world == 0) return v;
  uint64_t const block = v >> e.cyc_shift;
  return ((block * e.cyc_world + e.cyc_rank) << e.cyc_shift) + (v & ((uint64_t(1) << e.cyc_shift) - 1));
}

struct EnumPlan {
  EnumView view;
  bool projected;
};

static uint64_t *device_binomials() {
  static uint64_t *d = nullptr;
  if (d == nullptr) {
    init_binomials();
    CUDA_CHECK(cudaMalloc(&d, sizeof g_binom));
    CUDA_CHECK(cudaMemcpy(d, g_binom, sizeof g_binom, cudaMemcpyHostToDevice));
  }
  return d;
}

static EnumPlan make_plan(ls_hs_basis const *b, BasisInfo const &info) {
  EnumPlan plan{};
  EnumView &e = plan.view;
  e.number_bits = info.number_bits;
  e.binom = device_binomials();
  plan.projected = info.has_permutation_symmetries;
  uint64_t lo = info.min_state, hi = info.max_state;
  if (info.spinful_sectors) {
    int const n = b->number_sites;
    e.mode = 2;
    e.n_sites = n;
    e.hw = b->number_up;
    e.hw_b = b->number_particles - b->number_up;
    e.count_a = binomial(n, e.hw);
    e.total = e.count_a * binomial(n, e.hw_b);
    return plan;
  }
  if (!plan.projected && info.has_spin_inversion) {
    // StatesEnumeration.chpl:276-281
    uint64_t const mask = low_ones(b->number_sites);
    hi = std::min(hi, hi ^ mask);
  }
  e.min_state = lo;
  if (info.fixed_hamming) {
    e.mode = 1;
    e.hw = info.hamming_weight;
    e.rank_min = fixed_hamming_state_to_index(lo);
    e.total = fixed_hamming_state_to_index(hi) - e.rank_min + 1;
  } else {
    e.mode = 0;
    e.total = hi - lo + 1;  // number_bits == 64 without constraints is not enumerable anyway
  }
  return plan;
}

// ---- device-side candidate generation ----------------------------------------------
__device__ __forceinline__ uint64_t dev_binom(uint64_t const *__restrict__ binom, int n, int k) {
  return (k > n) ? 0 : __ldg(binom + n * 65 + k);
}
__device__ __forceinline__ uint64_t dev_unrank(uint64_t const *__restrict__ binom, uint64_t index,
                                               int hw, int number_bits) {
  uint64_t state = 0;
  int c = number_bits;
  for (int i = hw; i > 0; --i) {
    uint64_t b;
    do { --c; b = dev_binom(binom, c, i); } while (b > index);
    state |= uint64_t(1) << c;
    index -= b;
  }
  return state;
}

// Sequential generator of candidates k, k+1, ...
struct CandidateIter {
  uint64_t a, b;  // mode 2: a = up part, b = down part; else a = state
  uint64_t ka;    // mode 2: index of a within its sector
  __device__ __forceinline__ void init(EnumView const &e, uint64_t k) {
    if (e.mode == 0) {
      a = e.min_state + k;
    } else if (e.mode == 1) {
      a = dev_unrank(e.binom, e.rank_min + k, e.hw, e.number_bits);
    } else {
      uint64_t const kb = k / e.count_a;
      ka = k - kb * e.count_a;
      a = dev_unrank(e.binom, ka, e.hw, e.n_sites);
      b = dev_unrank(e.binom, kb, e.hw_b, e.n_sites);
    }
  }
  __device__ __forceinline__ uint64_t value(EnumView const &e) const {
    return e.mode == 2 ? ((b << e.n_sites) | a) : a;
  }
  __device__ __forceinline__ void next(EnumView const &e) {
    if (e.mode == 0) {
      a += 1;
    } else if (e.mode == 1) {
      // Gosper step; when the lowest run of ones does not reach bit 31 the step stays inside the low word
      uint32_t const l = (uint32_t)a, tl = l | (l - 1);
      if (l != 0 && tl != 0xffffffffu)
        a = (a & 0xffffffff00000000ull) | (uint64_t)((tl + 1) | (((~tl & (tl + 1)) - 1) >> __ffs((int)l)));
      else
        a = next_same_popcount(a);
    } else {
      if (++ka == e.count_a) {
        ka = 0;
        a = (e.hw == 0) ? 0 : ((uint64_t(1) << e.hw) - 1);
        b = (e.hw_b == 0) ? 0 : next_same_popcount(b);
      } else {
        a = next_same_popcount(a);
      }
    }
  }
};

// ---- unprojected enumeration ----------------------------------------------------------
__global__ void __launch_bounds__(256)
generate_states_kernel(EnumView e, uint64_t k_begin, uint64_t count, uint64_t *__restrict__ out) {
  uint64_t const chunk = 32;
  uint64_t const nchunks = (count + chunk - 1) / chunk;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks;
       c += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t const first = c * chunk;
    uint64_t last = min(first + chunk, count);
    uint64_t const k_actual = e.cyc_world == 0 ? k_begin + first : actual_word(e, c) * 32;  // (cyclic shares start at 0)
    if (e.cyc_world != 0) last = min(last, first + (e.total - k_actual));
    CandidateIter it;
    it.init(e, k_actual);
    for (uint64_t k = first; k < last; ++k) {
      out[k] = it.value(e);
      if (k + 1 < last) it.next(e);
    }
  }
}

// ---- pass A: survivor flags ---------------------------------------------------------------
constexpr int kBuildThreads = 128;

// Recompute the stabiliser character sum of lane `lane` with the scalar walk,
// in group order, exactly like the oracle.
template <int NP>
__device__ __noinline__ double lane_norm_sum(GroupView const &g, uint64_t const *smasks,
                                             uint32_t const (&xr)[NP], int lane) {
  uint64_t x = 0;
#pragma unroll
  for (int i = 0; i < NP; ++i) x |= (uint64_t)((xr[i] >> lane) & 1u) << i;
  uint64_t rep;
  double c_re, c_im, n;
  state_info_scalar<uint64_t>(g, smasks, x, rep, c_re, c_im, n);
  return n;
}

// One thread = 32 consecutive candidates (one "word").
//   alive_out[w]  bit k: candidate 32w+k is a representative with norm > 0
//   event_out[w]  bit k: ... and has a non-trivial stabiliser (norm != 1/sqrt|G|)
// FILTER: first pass of the two-phase build -- only the first `rows` table rows, only the
// "some image is smaller" verdict (no stabiliser events, no norms).  `list` != nullptr: the 32
// states of word w are list[32 w ..] (the survivors of the filter) instead of candidates by index.
template <int NP, bool INV, bool FILTER>
__global__ void __launch_bounds__(kBuildThreads, (NP <= 40 ? 5 : NP <= 52 ? 4 : 3))  // NP <= 40: five CTAs per SM (<= 102 registers)
build_flags_bitsliced_kernel(GroupView g, EnumView e, uint64_t word_begin, uint64_t number_words,
                             int identity_first, int rows, uint64_t const *__restrict__ list, uint64_t list_count,
                             uint32_t *__restrict__ alive_out, uint32_t *__restrict__ event_out,
                             uint32_t *__restrict__ block_counts) {
  extern __shared__ unsigned char smem_raw[];
  uint64_t *smasks = reinterpret_cast<uint64_t *>(smem_raw);
  uint32_t *planes = reinterpret_cast<uint32_t *>(smasks + (size_t)g.depth * g.number_masks);
  __shared__ uint32_t warp_counts[kBuildThreads / 32];
  if (!FILTER) stage_masks<uint64_t>(g, smasks);

  int const tid = threadIdx.x;
  uint64_t const w = (uint64_t)blockIdx.x * kBuildThreads + tid;
  uint32_t alive = 0, events = 0;
  uint32_t xr[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) xr[i] = 0;

  if (w < number_words) {
    uint64_t const k0 = (list != nullptr ? word_begin + w : actual_word(e, word_begin + w)) * 32;
    uint64_t const remaining = (list != nullptr ? list_count : e.total) - k0;
    int const valid = remaining >= 32 ? 32 : (int)remaining;
    alive = valid == 32 ? 0xffffffffu : ((1u << valid) - 1u);
    uint32_t lo[32], hi[32];
    if (list != nullptr) {
      uint64_t v = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < valid) v = __ldg(list + k0 + k);
        lo[k] = (uint32_t)v;
        hi[k] = (uint32_t)(v >> 32);
      }
    } else {
      CandidateIter it;
      it.init(e, k0);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        uint64_t const v = it.value(e);
        lo[k] = (uint32_t)v;
        hi[k] = (uint32_t)(v >> 32);
        if (k + 1 < valid) it.next(e);
      }
    }
    // the 32 states ascend (padding repeats the last one): equal ends mean equal everywhere, and a constant
    // bit is a constant plane -- the common case for everything above the low ~10 bits
    // (ascending as 64-bit values: the low words ascend only where the high words agree)
    bool const same_hi = NP <= 32 || hi[0] == hi[31];
    if (same_hi && ((lo[0] ^ lo[31]) >> 16) == 0) transpose32_low16(lo); else transpose32(lo);
    if (NP > 32) {
      if (same_hi) {
        uint32_t const h = hi[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) hi[i] = ((h >> i) & 1u) ? 0xffffffffu : 0u;
      } else {
        transpose32(hi);
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) xr[i] = (i < 32) ? lo[i] : hi[i - 32];
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) planes[i * kBuildThreads + tid] = xr[i];
  __syncwarp();  // a thread only ever reads its own column

  unsigned char const *column = reinterpret_cast<unsigned char const *>(planes + tid);
  int const G = min(rows, g.number_masks);
  int const nbits = g.number_bits;
#pragma unroll 1
  for (int j = 0; j < G; ++j) {
    PlaneRow<NP> const off(j);
    // With spin inversion both y and ~y must be >= x; the smaller of the two is
    // z = y ^ top(y) (top = the most significant live bit decides their order),
    // so one comparison per plane covers both images.  The table rows are sorted
    // by the source of that top bit (plane_table.cuh): the thread's column holds
    // the planes already XOR-ed with the current flip plane and is re-targeted in
    // place only when the class changes (every |G| / number_bits rows).
    if (INV) {
      uint32_t const ro = off[NP + 2];
      if (ro != kNoRetarget) {
        uint32_t const d = *reinterpret_cast<uint32_t const *>(column + ro);
#pragma unroll
        for (int i = 0; i < NP; ++i)
          if (i < NP - 3 || i < nbits) planes[i * kBuildThreads + tid] ^= d;  // padding planes stay zero
      }
    }
    uint32_t lt = 0, eq = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      uint32_t const z = *reinterpret_cast<uint32_t const *>(column + off[i]);
      if (FILTER) {
        uint32_t const d = z ^ xr[i];
        lt = (d & xr[i]) | (~d & lt);
      } else {
        cmp_step(z, xr[i], lt, eq);
      }
    }
    alive &= ~lt;
    if (!FILTER) events |= (identity_first && j == 0) ? 0u : eq;
    if (__all_sync(0xffffffffu, alive == 0)) break;
  }
  events &= alive;
  // Lanes with a non-trivial stabiliser: decide norm > 0 with the exact sum.
  uint32_t pending = FILTER ? 0u : events;
  while (pending != 0) {
    int const lane = __ffs((int)pending) - 1;
    pending &= pending - 1;
    double const s = lane_norm_sum<NP>(g, smasks, xr, lane);
    if (!(s > kNormThreshold)) {
      alive &= ~(1u << lane);
      events &= ~(1u << lane);
    }
  }
  if (w < number_words) {
    alive_out[w] = alive;
    if (!FILTER) event_out[w] = events;
  }
  // block survivor count
  unsigned c = (unsigned)__popc(alive);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((tid & 31) == 0) warp_counts[tid >> 5] = c;
  __syncthreads();
  if (tid == 0) {
    unsigned total = 0;
#pragma unroll
    for (int i = 0; i < kBuildThreads / 32; ++i) total += warp_counts[i];
    block_counts[blockIdx.x] = total;
  }
}

// Scalar variant of pass A (Benes walk per candidate, early exit); used when
// the permutation table does not fit constant memory and for A/B validation
// (LS_B200_BUILD=scalar).
__global__ void __launch_bounds__(kBuildThreads)
build_flags_scalar_kernel(GroupView g, EnumView e, uint64_t word_begin, uint64_t number_words,
                          uint32_t *__restrict__ alive_out, uint32_t *__restrict__ event_out,
                          uint32_t *__restrict__ block_counts) {
  extern __shared__ unsigned char smem_raw[];
  uint64_t *smasks = reinterpret_cast<uint64_t *>(smem_raw);
  __shared__ uint32_t warp_counts[kBuildThreads / 32];
  stage_masks<uint64_t>(g, smasks);
  OrbitScalar<uint64_t> orbit{g, smasks};
  int const tid = threadIdx.x;
  uint64_t const w = (uint64_t)blockIdx.x * kBuildThreads + tid;
  uint32_t alive = 0, events = 0;
  if (w < number_words) {
    uint64_t const k0 = actual_word(e, word_begin + w) * 32;
    uint64_t const remaining = e.total - k0;
    int const valid = remaining >= 32 ? 32 : (int)remaining;
    CandidateIter it;
    it.init(e, k0);
    for (int k = 0; k < valid; ++k) {
      uint64_t const x = it.value(e);
      bool flag = true;
      double s = 0.0;
      for (int j = 0; j < g.number_masks; ++j) {
        uint64_t const y = orbit.image(x, j);
        if (y < x) { flag = false; break; }
        if (y == x) s += g.re[j];
        if (g.spin_inversion != 0) {
          uint64_t const yf = y ^ g.flip_mask;
          if (yf < x) { flag = false; break; }
          if (yf == x) s += (double)g.spin_inversion * g.re[j];
        }
      }
      if (flag && s > kNormThreshold) {
        alive |= 1u << k;
        if (s != 1.0) events |= 1u << k;
      }
      if (k + 1 < valid) it.next(e);
    }
    alive_out[w] = alive;
    event_out[w] = events;
  }
  unsigned c = (unsigned)__popc(alive);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((tid & 31) == 0) warp_counts[tid >> 5] = c;
  __syncthreads();
  if (tid == 0) {
    unsigned total = 0;
    for (int i = 0; i < kBuildThreads / 32; ++i) total += warp_counts[i];
    block_counts[blockIdx.x] = total;
  }
}

// ---- pass C: ordered scatter ------------------------------------------------------------------
__global__ void __launch_bounds__(kBuildThreads)
build_scatter_kernel(GroupView g, EnumView e, uint64_t word_begin, uint64_t number_words,
                     uint64_t const *__restrict__ list, uint32_t const *__restrict__ alive_in,
                     uint32_t const *__restrict__ event_in, uint32_t const *__restrict__ block_offsets,
                     uint64_t out_base, uint64_t *__restrict__ reps_out, double *__restrict__ norms_out) {
  extern __shared__ unsigned char smem_raw[];
  uint64_t *smasks = reinterpret_cast<uint64_t *>(smem_raw);
  __shared__ uint32_t warp_sums[kBuildThreads / 32];
  // norms_out == nullptr: first phase of the two-phase build, only the surviving states are wanted
  if (norms_out != nullptr) stage_masks<uint64_t>(g, smasks);
  int const tid = threadIdx.x;
  uint64_t const w = (uint64_t)blockIdx.x * kBuildThreads + tid;
  uint32_t const alive = (w < number_words) ? alive_in[w] : 0u;
  uint32_t const events = (w < number_words && event_in != nullptr) ? event_in[w] : 0u;
  // exclusive scan of popcounts within the block
  unsigned const mine = (unsigned)__popc(alive);
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned const t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  unsigned warp_base = 0;
  for (int i = 0; i < (tid >> 5); ++i) warp_base += warp_sums[i];
  if (alive == 0) return;
  uint64_t pos = out_base + (uint64_t)block_offsets[blockIdx.x] + warp_base + (incl - mine);
  double const trivial_norm = norm_from_sum(g, 1.0);
  uint64_t const k0 = (list != nullptr ? word_begin + w : actual_word(e, word_begin + w)) * 32;
  int const last = 31 - __clz((int)alive);
  CandidateIter it;
  if (list == nullptr) it.init(e, k0);
  for (int k = 0; k <= last; ++k) {
    if ((alive >> k) & 1u) {
      uint64_t const x = list != nullptr ? __ldg(list + k0 + k) : it.value(e);
      double norm = trivial_norm;
      if ((events >> k) & 1u) {
        uint64_t rep;
        double c_re, c_im, n;
        state_info_scalar<uint64_t>(g, smasks, x, rep, c_re, c_im, n);
        norm = norm_from_sum(g, n);
      }
      reps_out[pos] = x;
      if (norms_out != nullptr) norms_out[pos] = norm;
      ++pos;
    }
    if (k < last && list == nullptr) it.next(e);
  }
}

// ---- host driver ---------------------------------------------------------------------------------
using FlagsKernel = void (*)(GroupView, EnumView, uint64_t, uint64_t, int, int, uint64_t const *, uint64_t, uint32_t *,
                             uint32_t *, uint32_t *);

template <int NP>
static FlagsKernel pick_inv(bool inv, bool filter) {
  if (filter) return inv ? build_flags_bitsliced_kernel<NP, true, true> : build_flags_bitsliced_kernel<NP, false, true>;
  return inv ? build_flags_bitsliced_kernel<NP, true, false> : build_flags_bitsliced_kernel<NP, false, false>;
}
static FlagsKernel pick_flags_kernel(int np, bool inv, bool filter) {
  switch (np) {
    case 4: return pick_inv<4>(inv, filter);
    case 8: return pick_inv<8>(inv, filter);
    case 12: return pick_inv<12>(inv, filter);
    case 16: return pick_inv<16>(inv, filter);
    case 20: return pick_inv<20>(inv, filter);
    case 24: return pick_inv<24>(inv, filter);
    case 28: return pick_inv<28>(inv, filter);
    case 32: return pick_inv<32>(inv, filter);
    case 36: return pick_inv<36>(inv, filter);
    case 40: return pick_inv<40>(inv, filter);
    case 44: return pick_inv<44>(inv, filter);
    case 48: return pick_inv<48>(inv, filter);
    case 52: return pick_inv<52>(inv, filter);
    case 56: return pick_inv<56>(inv, filter);
    case 60: return pick_inv<60>(inv, filter);
    case 64: return pick_inv<64>(inv, filter);
  }
  return nullptr;
}

// Build the representatives whose candidate index lies in one of the (ascending, disjoint)
// ranges; the output is their concatenation, counts[i] the number found in range i.  One call
// serves a rank's whole block-cyclic share: scratch and output are allocated once.
static bool want_managed_view();
// first[i] = position in the (sorted) output of the first representative that is >= the first candidate of the
// rank's i-th block; first[number_blocks] = count
__global__ void __launch_bounds__(256)
cyclic_block_starts_kernel(EnumView e, uint64_t number_blocks, uint64_t const *__restrict__ reps, uint64_t count,
                           uint64_t *__restrict__ first) {
  uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > number_blocks) return;
  if (i == number_blocks) {
    first[i] = count;
    return;
  }
  uint64_t const k = actual_word(e, i << e.cyc_shift) * 32;
  CandidateIter it;
  it.init(e, k);
  uint64_t const state = it.value(e);
  uint64_t lo = 0, hi = count;
  while (lo < hi) {
    uint64_t const mid = lo + (hi - lo) / 2;
    if (__ldg(reps + mid) < state) lo = mid + 1;
    else hi = mid;
  }
  first[i] = lo;
}

CyclicShare cyclic_share(uint64_t total, int block_shift, int world, int rank) {
  CyclicShare c;
  c.shift = block_shift;
  c.world = world;
  c.rank = rank;
  uint64_t const block = uint64_t(32) << block_shift;  // candidates per block
  uint64_t const blocks_total = (total + block - 1) / block;
  c.blocks_total = blocks_total;
  c.number_blocks = (uint64_t)rank < blocks_total ? (blocks_total - (uint64_t)rank + (uint64_t)world - 1) / (uint64_t)world : 0;
  c.virtual_candidates = 0;
  if (c.number_blocks > 0) {
    uint64_t const last_block = (c.number_blocks - 1) * (uint64_t)world + (uint64_t)rank;
    uint64_t const last_begin = last_block * block;
    c.virtual_candidates = (c.number_blocks - 1) * block + (std::min(total, last_begin + block) - last_begin);
  }
  return c;
}

BuildResult build_ranges(ls_hs_basis const *basis, Ranges ranges, std::vector<uint64_t> *counts, bool host_visible,
                         CyclicShare const *cyc) {
  Runtime &rt = runtime();
  BasisInfo const info = basis_info(basis);
  LSB_CHECK(info.number_bits <= 64, "bases with more than 64 bits are not supported");
  EnumPlan plan = make_plan(basis, info);
  EnumView &e = plan.view;
  if (cyc != nullptr) {
    // a rank's block-cyclic share (dist.cu) scanned as ONE virtual range: the kernels translate word numbers, so the
    // launches are as large as those of a single-GPU build; counts = representatives per block (from the output)
    e.cyc_shift = cyc->shift;
    e.cyc_world = (uint32_t)cyc->world;
    e.cyc_rank = (uint32_t)cyc->rank;
    ranges = Ranges{{0, cyc->virtual_candidates}};
  }
  uint64_t candidates = 0, longest = 0;
  for (auto &r : ranges) {
    if (cyc == nullptr) r.second = std::min(r.second, e.total);
    r.first = std::min(r.first, r.second);
    candidates += r.second - r.first;
    longest = std::max(longest, r.second - r.first);
  }
  if (counts != nullptr) counts->assign(cyc != nullptr ? (size_t)cyc->number_blocks : ranges.size(), 0);
  BuildResult res;
  if (candidates == 0) return res;
  static bool const profile = getenv("LS_B200_PROFILE") != nullptr;
  auto const wall0 = std::chrono::steady_clock::now();
  double alloc_ms = 0, sync_ms = 0;
  auto since = [](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  };
  CUDA_CHECK(cudaEventRecord(rt.ev0, rt.stream));

  // host_visible: the representatives go straight into the managed allocation that will serve as the caller's
  // host view (make_host_view then has nothing to copy)
  auto alloc_reps = [&](uint64_t **p, uint64_t n) {
    *p = nullptr;
    if (host_visible && want_managed_view()) *p = static_cast<uint64_t *>(block_alloc(sizeof(uint64_t) * n, true));
    if (*p == nullptr) *p = static_cast<uint64_t *>(block_alloc(sizeof(uint64_t) * n, false));
  };
  if (!plan.projected) {
    alloc_reps(&res.d_reps, candidates);
    for (size_t i = 0; i < ranges.size(); ++i) {
      uint64_t const count = ranges[i].second - ranges[i].first;
      if (count == 0) continue;
      unsigned const blocks = (unsigned)std::min<uint64_t>((count / 32 + 256) / 256, (uint64_t)rt.sm_count * 16);
      generate_states_kernel<<<blocks, 256, 0, rt.stream>>>(e, ranges[i].first, count, res.d_reps + res.count);
      count_launch();
      CUDA_CHECK(cudaGetLastError());
      res.count += count;
      if (counts != nullptr && cyc == nullptr) (*counts)[i] = count;
    }
  } else {
    GroupData const &g = *info.group;
    int const np = std::max(4, (g.number_bits + 3) / 4 * 4);
    bool const inv = g.spin_inversion != 0;
    char const *mode = getenv("LS_B200_BUILD");
    bool use_scalar = mode != nullptr && strcmp(mode, "scalar") == 0;
    size_t const masks_bytes = sizeof(uint64_t) * (size_t)g.depth * (size_t)g.number_masks;
    size_t const smem_bitsliced = masks_bytes + (size_t)np * kBuildThreads * 4;
    if (!use_scalar && (smem_bitsliced > rt.smem_optin || !upload_plane_offsets(g, np, kBuildThreads))) use_scalar = true;
    LSB_CHECK(masks_bytes <= rt.smem_optin, "symmetry group too large for shared memory staging");
    bool const identity_first = identity_is_first(g);
    FlagsKernel flags_kernel = use_scalar ? nullptr : pick_flags_kernel(np, inv, false);
    if (flags_kernel != nullptr && smem_bitsliced > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(flags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bitsliced));
    // Two-phase build: a candidate survives the first few group elements with probability ~1/(2 K + 1), so
    // the full orbit walk (and its stabiliser bookkeeping) runs only on the compacted survivors of a cheap
    // K-row filter.  LS_B200_BUILD=onepass keeps the single full pass; LS_B200_BUILD_FILTER_ROWS sets K.
    int filter_rows = 8;
    if (char const *env = getenv("LS_B200_BUILD_FILTER_ROWS")) filter_rows = std::max(1, atoi(env));
    bool const two_phase = !use_scalar && g.number_masks >= 4 * filter_rows && !(mode != nullptr && strcmp(mode, "onepass") == 0);
    FlagsKernel filter_kernel = two_phase ? pick_flags_kernel(np, inv, true) : nullptr;
    if (filter_kernel != nullptr && smem_bitsliced > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bitsliced));

    if (masks_bytes > 48 * 1024) {
      CUDA_CHECK(cudaFuncSetAttribute(build_flags_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)masks_bytes));
      CUDA_CHECK(cudaFuncSetAttribute(build_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)masks_bytes));
    }

    // words per super-chunk (2^28 candidates): scratch stays at ~64 MB + survivors -- cudaMalloc / cudaFree of
    // GB-sized buffers showed 2..170 ms run-to-run variance on the test boxes, more than the kernels
    uint64_t const super = uint64_t(1) << 23;
    uint64_t const super_words = std::min((longest + 31) / 32, super);
    uint64_t const super_blocks = (super_words + kBuildThreads - 1) / kBuildThreads;
    // scratch lives for the process (grow-only): cudaFree of these buffers at the end of every build cost
    // 10..200 ms on the test boxes (device-wide synchronisation + unmapping), far more than the kernels
    static DeviceBuffer<uint32_t> alive, events, block_counts, block_offsets;
    static DeviceBuffer<unsigned char> scan_tmp;
    static DeviceBuffer<uint64_t> survivors;
    auto const t_alloc = std::chrono::steady_clock::now();
    alive.reserve(super_words);
    events.reserve(super_words);
    block_counts.reserve(super_blocks + 1);
    block_offsets.reserve(super_blocks + 1);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, block_counts.ptr, block_offsets.ptr, (int)(super_blocks + 1), rt.stream);
    scan_tmp.reserve(tmp_bytes);

    // Output capacity: orbit-counting estimate, grown on demand.
    // (with spin inversion the candidate range is already halved -- top bit clear, Basis.hs:736-740 -- so the
    // orbits cover it |G| times, not 2 |G| times)
    uint64_t const images = (uint64_t)std::max(1, g.number_masks);
    uint64_t capacity = std::min<uint64_t>(candidates, candidates / images * 5 / 4 + (1u << 16));
    alloc_reps(&res.d_reps, capacity);
    block_alloc(&res.d_norms, sizeof(double) * capacity);
    alloc_ms += since(t_alloc);
    uint64_t emitted = 0, scanned = 0;
    GroupView const gv = g.view();

    for (size_t i = 0; i < ranges.size(); ++i) {
      uint64_t const k_begin = ranges[i].first, k_end = ranges[i].second;
      if (k_begin >= k_end) continue;
      LSB_CHECK(k_begin % 32 == 0, "shard boundaries must be multiples of 32 candidates");
      uint64_t const words_total = (k_end - k_begin + 31) / 32;
      uint64_t const word0 = k_begin / 32;
      uint64_t const emitted_before = emitted;
      for (uint64_t done = 0; done < words_total; done += super) {
        uint64_t const nwords = std::min(super, words_total - done);
        unsigned const blocks = (unsigned)((nwords + kBuildThreads - 1) / kBuildThreads);
        // Clip the enumeration so that the tail word of this range is masked.
        EnumView ev = e;
        if (cyc == nullptr) ev.total = k_end;
        // `source` / `source_words` / `source_blocks`: what the final scatter reads -- candidates by index, or
        // (two-phase) the compacted survivors of the filter
        uint64_t const *source = nullptr;
        uint64_t source_words = nwords;
        unsigned source_blocks = blocks;
        uint64_t source_word0 = word0 + done;
        auto scan_counts = [&](unsigned nblocks) -> uint32_t {
          CUDA_CHECK(cudaMemsetAsync(block_counts.ptr + nblocks, 0, sizeof(uint32_t), rt.stream));
          cub::DeviceScan::ExclusiveSum(scan_tmp.ptr, tmp_bytes, block_counts.ptr, block_offsets.ptr, (int)(nblocks + 1), rt.stream);
          count_launch();
          uint32_t total = 0;
          auto const t_sync = std::chrono::steady_clock::now();
          CUDA_CHECK(cudaMemcpyAsync(&total, block_offsets.ptr + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, rt.stream));
          CUDA_CHECK(cudaStreamSynchronize(rt.stream));
          sync_ms += since(t_sync);
          return total;
        };
        uint32_t chunk_total = 0;
        if (use_scalar) {
          build_flags_scalar_kernel<<<blocks, kBuildThreads, masks_bytes, rt.stream>>>(
              gv, ev, word0 + done, nwords, alive.ptr, events.ptr, block_counts.ptr);
          count_launch();
          CUDA_CHECK(cudaGetLastError());
          chunk_total = scan_counts(blocks);
        } else if (!two_phase) {
          flags_kernel<<<blocks, kBuildThreads, smem_bitsliced, rt.stream>>>(
              gv, ev, word0 + done, nwords, identity_first ? 1 : 0, g.number_masks, nullptr, 0, alive.ptr, events.ptr,
              block_counts.ptr);
          count_launch();
          CUDA_CHECK(cudaGetLastError());
          chunk_total = scan_counts(blocks);
        } else {
          // phase 1: K-row filter over the candidates, survivors compacted (in order) into a list
          filter_kernel<<<blocks, kBuildThreads, smem_bitsliced, rt.stream>>>(
              gv, ev, word0 + done, nwords, identity_first ? 1 : 0, filter_rows, nullptr, 0, alive.ptr, events.ptr,
              block_counts.ptr);
          count_launch();
          CUDA_CHECK(cudaGetLastError());
          uint32_t const alive_total = scan_counts(blocks);
          if (alive_total > 0) {
            auto const t_grow = std::chrono::steady_clock::now();
            survivors.reserve((size_t)alive_total + 32);
            alloc_ms += since(t_grow);
            build_scatter_kernel<<<blocks, kBuildThreads, masks_bytes, rt.stream>>>(
                gv, ev, word0 + done, nwords, nullptr, alive.ptr, nullptr, block_offsets.ptr, 0, survivors.ptr, nullptr);
            count_launch();
            // phase 2: the full walk on the survivors
            source = survivors.ptr;
            source_words = ((uint64_t)alive_total + 31) / 32;
            source_blocks = (unsigned)((source_words + kBuildThreads - 1) / kBuildThreads);
            source_word0 = 0;
            flags_kernel<<<source_blocks, kBuildThreads, smem_bitsliced, rt.stream>>>(
                gv, ev, 0, source_words, identity_first ? 1 : 0, g.number_masks, survivors.ptr, alive_total, alive.ptr,
                events.ptr, block_counts.ptr);
            count_launch();
            CUDA_CHECK(cudaGetLastError());
            chunk_total = scan_counts(source_blocks);
          }
        }
        scanned += std::min<uint64_t>(nwords * 32, k_end - k_begin - done * 32);
        if (emitted + chunk_total > capacity) {
          // density so far, with head room; never more than what is left to scan
          uint64_t const remaining = candidates - scanned;
          uint64_t const projected =
              emitted + chunk_total + std::min<uint64_t>(remaining, (uint64_t)((double)(emitted + chunk_total) / (double)scanned * (double)remaining * 1.1) + (1u << 16));
          uint64_t const new_capacity = std::max<uint64_t>(emitted + chunk_total, projected);
          uint64_t *new_reps = nullptr;
          double *new_norms = nullptr;
          auto const t_grow = std::chrono::steady_clock::now();
          alloc_reps(&new_reps, new_capacity);
          block_alloc(&new_norms, sizeof(double) * new_capacity);
          CUDA_CHECK(cudaMemcpyAsync(new_reps, res.d_reps, sizeof(uint64_t) * emitted, cudaMemcpyDeviceToDevice, rt.stream));
          CUDA_CHECK(cudaMemcpyAsync(new_norms, res.d_norms, sizeof(double) * emitted, cudaMemcpyDeviceToDevice, rt.stream));
          CUDA_CHECK(cudaStreamSynchronize(rt.stream));
          block_free(res.d_reps);
          block_free(res.d_norms);
          res.d_reps = new_reps;
          res.d_norms = new_norms;
          capacity = new_capacity;
          alloc_ms += since(t_grow);
        }
        if (chunk_total > 0) {
          build_scatter_kernel<<<source_blocks, kBuildThreads, masks_bytes, rt.stream>>>(
              gv, ev, source_word0, source_words, source, alive.ptr, events.ptr, block_offsets.ptr, emitted, res.d_reps,
              res.d_norms);
          count_launch();
          CUDA_CHECK(cudaGetLastError());
        }
        emitted += chunk_total;
      }
      if (counts != nullptr && cyc == nullptr) (*counts)[i] = emitted - emitted_before;
    }
    res.count = emitted;
  }
  if (cyc != nullptr && counts != nullptr && cyc->number_blocks > 0) {
    static DeviceBuffer<uint64_t> starts;
    uint64_t const nb = cyc->number_blocks;
    uint64_t *d = starts.reserve((size_t)nb + 1);
    cyclic_block_starts_kernel<<<(unsigned)((nb + 256) / 256), 256, 0, rt.stream>>>(e, nb, res.d_reps, res.count, d);
    count_launch();
    CUDA_CHECK(cudaGetLastError());
    std::vector<uint64_t> h((size_t)nb + 1);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d, sizeof(uint64_t) * (nb + 1), cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    for (uint64_t b = 0; b < nb; ++b) (*counts)[(size_t)b] = h[(size_t)b + 1] - h[(size_t)b];
    res.d_block_starts = d;
  }
  CUDA_CHECK(cudaEventRecord(rt.ev1, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, rt.ev0, rt.ev1));
  rt.last_build_ms = ms;
  if (profile)
    fprintf(stderr, "[ls_b200] build_ranges: %llu candidates -> %llu states; device span %.2f ms, wall %.2f ms of which "
                    "allocations %.2f ms, waiting for kernels + counts %.2f ms\n",
            (unsigned long long)candidates, (unsigned long long)res.count, ms, since(wall0), alloc_ms, sync_ms);
  return res;
}

static BuildResult build_range(ls_hs_basis const *basis, uint64_t k_begin, uint64_t k_end, bool host_visible = false) {
  return build_ranges(basis, Ranges{{k_begin, k_end}}, nullptr, host_visible);
}

static void free_pinned(void *p) {
  if (p == nullptr) return;
  {
    std::lock_guard<std::mutex> lock(runtime().mutex);
    auto &reg = built_registry();
    auto it = reg.find(p);
    if (it != reg.end()) {
      if (it->second.d_reps != p) block_free(it->second.d_reps);
      block_free(it->second.d_norms);
      reg.erase(it);
    }
  }
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeManaged) block_free(p);
  else cudaFreeHost(p);
}

// The reference hands the representatives to its callers as a HOST array
// (chpl_external_array::elts, read directly by Python and Haskell).  Default:
// one managed allocation serves as both the device array of the kernels and
// that host view -- resident in HBM, advised read-mostly, so a host read
// faults in a read-only duplicate of the touched pages and the device copy
// stays put.  Nothing is copied at build time.  LS_B200_HOST_MIRROR=pinned
// keeps a second, pinned host copy instead (cudaMallocHost of the whole array:
// ~0.4 ms per MB, more than the build itself for large bases).
// Takes ownership of d_reps; returns the host-visible pointer and updates d_reps.
static bool want_managed_view() {
  char const *mode = getenv("LS_B200_HOST_MIRROR");
  if (mode != nullptr && strcmp(mode, "pinned") == 0) return false;
  int concurrent = 0;
  cudaDeviceGetAttribute(&concurrent, cudaDevAttrConcurrentManagedAccess, runtime().device);
  return concurrent != 0;
}
static uint64_t *make_host_view(uint64_t *&d_reps, uint64_t count) {
  Runtime &rt = runtime();
  size_t const bytes = sizeof(uint64_t) * count;
  bool const managed = want_managed_view();
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, d_reps) == cudaSuccess && attr.type == cudaMemoryTypeManaged) {
    // built in place (build_ranges, host_visible): only the advice is missing
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    CUDA_CHECK(cudaMemAdvise(d_reps, bytes, cudaMemAdviseSetReadMostly, rt.device));
    return d_reps;
  }
  (void)cudaGetLastError();
  if (managed) {
    uint64_t *m = static_cast<uint64_t *>(block_alloc(bytes, true));
    if (m != nullptr) {
      CUDA_CHECK(cudaMemcpyAsync(m, d_reps, bytes, cudaMemcpyDeviceToDevice, rt.stream));
      CUDA_CHECK(cudaStreamSynchronize(rt.stream));
      CUDA_CHECK(cudaMemAdvise(m, bytes, cudaMemAdviseSetReadMostly, rt.device));
      block_free(d_reps);
      d_reps = m;
      return m;
    }
  }
  uint64_t *host = nullptr;
  CUDA_CHECK(cudaMallocHost(&host, bytes));
  CUDA_CHECK(cudaMemcpyAsync(host, d_reps, bytes, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  return host;
}

IndexData *create_index(uint64_t const *host_reps, int64_t count, int number_bits, int prefix_bits);
void launch_state_info(GroupData const &g, int64_t n, uint64_t const *d_alphas, uint64_t *d_betas,
                       double2 *d_chars, double *d_norms);

// Norms of the representatives, computed lazily when the basis was populated
// through ls_hs_unchecked_set_representatives (the reference recomputes them on
// every matvec, BatchedOperator.chpl:222-238).
void ensure_norms(IndexData &ix, GroupData const &g) {
  if (ix.d_norms != nullptr || ix.number_states == 0) return;
  Runtime &rt = runtime();
  block_alloc(&ix.d_norms, sizeof(double) * (size_t)ix.number_states);
  DeviceBuffer<uint64_t> betas;
  DeviceBuffer<double2> chars;
  int64_t const chunk = int64_t(1) << 24;
  betas.reserve((size_t)std::min<int64_t>(chunk, ix.number_states));
  chars.reserve((size_t)std::min<int64_t>(chunk, ix.number_states));
  for (int64_t b = 0; b < ix.number_states; b += chunk) {
    int64_t const n = std::min(chunk, ix.number_states - b);
    launch_state_info(g, n, ix.d_reps + b, betas.ptr, chars.ptr, ix.d_norms + b);
  }
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
}

uint64_t *alloc_representatives(uint64_t count) {
  size_t const bytes = sizeof(uint64_t) * std::max<uint64_t>(count, 1);
  uint64_t *p = nullptr;
  if (getenv("LS_B200_NO_HOST_MIRROR") == nullptr && want_managed_view()) p = static_cast<uint64_t *>(block_alloc(bytes, true));
  if (p == nullptr) p = static_cast<uint64_t *>(block_alloc(bytes, false));
  return p;
}

uint64_t number_candidates(ls_hs_basis const *basis) { return make_plan(basis, basis_info(basis)).view.total; }

// Installs device-resident representatives (+ norms) as the basis' list; the library takes ownership of both
// buffers.  basis->representatives.elts (a host pointer in the reference) is served by a managed allocation that
// doubles as the device array, unless LS_B200_NO_HOST_MIRROR is set (then elts stays NULL: device-only basis).
void install_representatives(ls_hs_basis *basis, uint64_t *representatives_dev, double *norms_dev, uint64_t count,
                             int cache_bits) {
  uint64_t *host = nullptr;
  bool const mirror = getenv("LS_B200_NO_HOST_MIRROR") == nullptr;
  if (mirror && count > 0) host = make_host_view(representatives_dev, count);
  void const *key = host != nullptr ? (void const *)host : (void const *)representatives_dev;
  built_registry()[key] = BuiltReps{representatives_dev, norms_dev, count};
  basis->representatives.elts = host;
  basis->representatives.num_elts = count;
  basis->representatives.freer = host != nullptr ? reinterpret_cast<void *>(&free_pinned) : nullptr;
  int const number_bits = (basis->particle_type == LS_HS_SPINFUL_FERMION ? 2 : 1) * basis->number_sites;
  IndexData *ix = create_index(static_cast<uint64_t const *>(key), (int64_t)count, number_bits, cache_bits);
  ix->host_reps = host;
  if (host == nullptr) ix->owns_d_reps = true;  // no host view whose freer would release the device array
  ix->identity = basis->state_index_is_identity;
  basis->kernels->state_index_data = ix;
  basis->kernels->state_index_kernel = &ls_hs_state_index_binary_search_kernel;
}

}  // namespace lsb

using namespace lsb;

extern "C" {

int ls_b200_build_blocks(ls_hs_basis const *basis, uint64_t first_begin, uint64_t block_size, uint64_t stride,
                         uint64_t number_blocks, uint64_t **representatives_dev, double **norms_dev,
                         uint64_t *block_counts) {
  int status = -1;
  guarded(__func__, [&] {
    auto const t0 = std::chrono::steady_clock::now();
    Ranges ranges;
    for (uint64_t b = 0; b < number_blocks; ++b)
      ranges.emplace_back(first_begin + b * stride, first_begin + b * stride + block_size);
    std::vector<uint64_t> counts;
    BuildResult r = build_ranges(basis, ranges, &counts);
    for (uint64_t b = 0; b < number_blocks; ++b) block_counts[b] = counts[(size_t)b];
    if (getenv("LS_B200_PROFILE") != nullptr)
      fprintf(stderr, "[ls_b200] build_blocks %llu x [%llu + k %llu, +%llu): %llu states, device %.2f ms, wall %.2f ms\n",
              (unsigned long long)number_blocks, (unsigned long long)first_begin, (unsigned long long)stride,
              (unsigned long long)block_size, (unsigned long long)r.count, runtime().last_build_ms,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    *representatives_dev = r.d_reps;
    if (norms_dev != nullptr) *norms_dev = r.d_norms; else block_free(r.d_norms);
    status = 0;
  });
  return status;
}

uint64_t ls_b200_number_candidates(ls_hs_basis const *basis) {
  uint64_t total = 0;
  guarded(__func__, [&] { total = make_plan(basis, basis_info(basis)).view.total; });
  return total;
}

int ls_b200_build_shard(ls_hs_basis const *basis, uint64_t index_begin, uint64_t index_end,
                        uint64_t **representatives_dev, double **norms_dev, uint64_t *count) {
  int status = -1;
  guarded(__func__, [&] {
    auto const t0 = std::chrono::steady_clock::now();
    BuildResult r = build_range(basis, index_begin, index_end);
    if (getenv("LS_B200_PROFILE") != nullptr)
      fprintf(stderr, "[ls_b200] build_shard [%llu, %llu): %llu states, device %.2f ms, wall %.2f ms\n",
              (unsigned long long)index_begin, (unsigned long long)index_end, (unsigned long long)r.count,
              runtime().last_build_ms,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    *representatives_dev = r.d_reps;
    if (norms_dev != nullptr) *norms_dev = r.d_norms; else block_free(r.d_norms);
    *count = r.count;
    status = 0;
  });
  return status;
}

// chapel/src/StatesEnumeration.chpl:692-709.  `lower`/`upper` are ignored, as
// in the reference (:683-688 recompute them from the basis).
void ls_chpl_enumerate_representatives(ls_hs_basis const *basis, uint64_t lower, uint64_t upper,
                                       chpl_external_array *dest) {
  (void)lower;
  (void)upper;
  dest->elts = nullptr;
  dest->num_elts = 0;
  dest->freer = nullptr;
  guarded(__func__, [&] {
    static bool const profile = getenv("LS_B200_PROFILE") != nullptr;
    auto const t0 = std::chrono::steady_clock::now();
    BuildResult r = build_range(basis, 0, ~uint64_t(0), true);
    auto const t1 = std::chrono::steady_clock::now();
    uint64_t *host = nullptr;
    if (r.count > 0) {
      host = make_host_view(r.d_reps, r.count);
      auto const t2 = std::chrono::steady_clock::now();
      if (profile) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[ls_b200] enumerate: build_range %.2f ms, host view %.2f ms (%zu states, %s)\n", ms(t0, t1),
                ms(t1, t2), (size_t)r.count, host == r.d_reps ? "managed" : "pinned copy");
      }
      built_registry()[host] = BuiltReps{r.d_reps, r.d_norms, r.count};
    } else {
      block_free(r.d_reps);
      block_free(r.d_norms);
    }
    dest->elts = host;
    dest->num_elts = r.count;
    dest->freer = reinterpret_cast<void *>(&free_pinned);
  });
}

// kernels/reference.c:173-194
void ls_hs_build_representatives(ls_hs_basis *basis, uint64_t const lower, uint64_t const upper) {
  ls_chpl_kernels const *kernels = ls_hs_internal_get_chpl_kernels();
  LSB_CHECK(kernels->enumerate_states != nullptr,
            "enumerate_states kernel is NULL, ls_chpl_init was supposed to initialize it");
  if (basis->representatives.num_elts > 0) return;  // already built
  if (comm_world() > 1) {
    // A communicator is active (ls_b200_comm_init): the reference's multi-locale build
    // (chapel/src/StatesEnumeration.chpl:537-602) -- every rank scans its share of the candidates and keeps a
    // contiguous range of the sorted representatives; basis->representatives is this rank's block.
    if (index_of(basis) != nullptr) return;  // (a rank may own zero rows)
    guarded(__func__, [&] { dist_build_local(basis, nullptr, 0); });
    return;
  }
  auto const t0 = std::chrono::steady_clock::now();
  (*kernels->enumerate_states)(basis, lower, upper, &basis->representatives);
  // A failed enumeration (reported through ls_hs_error) leaves no freer behind: do not wire an index over the
  // empty array, so that the basis stays "not built" and the call can be retried.
  if (basis->representatives.freer == nullptr) return;
  auto const t1 = std::chrono::steady_clock::now();
  int const number_bits = (basis->particle_type == LS_HS_SPINFUL_FERMION ? 2 : 1) * basis->number_sites;
  int const default_cache_bits = 22;
  auto *ix = reinterpret_cast<IndexData *>(ls_hs_create_state_index_binary_search_kernel_data(
      &basis->representatives, number_bits, default_cache_bits));
  if (ix != nullptr) ix->identity = basis->state_index_is_identity;
  basis->kernels->state_index_data = ix;
  basis->kernels->state_index_kernel = &ls_hs_state_index_binary_search_kernel;
  if (getenv("LS_B200_PROFILE") != nullptr) {
    auto const t2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[ls_b200] build_representatives: enumerate %.2f ms, index %.2f ms\n", ms(t0, t1), ms(t1, t2));
  }
}

// kernels/reference.c:196-211 (borrows the caller's array)
void ls_hs_unchecked_set_representatives(ls_hs_basis *basis, chpl_external_array const *states,
                                         int const cache_bits) {
  LSB_CHECK(basis->representatives.num_elts == 0, "representatives have already been set");
  LSB_CHECK(basis->kernels != nullptr, "basis->kernels not set");
  LSB_CHECK(basis->kernels->state_index_kernel == nullptr, "state_index_kernel has already been set");
  basis->representatives = *states;
  int const number_bits = (basis->particle_type == LS_HS_SPINFUL_FERMION ? 2 : 1) * basis->number_sites;
  auto *ix = reinterpret_cast<IndexData *>(ls_hs_create_state_index_binary_search_kernel_data(
      &basis->representatives, number_bits, cache_bits));
  if (ix != nullptr) ix->identity = basis->state_index_is_identity;
  basis->kernels->state_index_data = ix;
  basis->kernels->state_index_kernel = &ls_hs_state_index_binary_search_kernel;
}

int ls_b200_set_representatives_device(ls_hs_basis *basis, uint64_t *representatives_dev,
                                       double *norms_dev, uint64_t count, int cache_bits) {
  int status = -1;
  LSB_CHECK(basis->representatives.num_elts == 0, "representatives have already been set");
  guarded(__func__, [&] {
    install_representatives(basis, representatives_dev, norms_dev, count, cache_bits);
    status = 0;
  });
  return status;
}

int ls_b200_basis_device_view(ls_hs_basis const *basis, uint64_t const **representatives,
                              double const **norms, uint64_t *count) {
  int status = -1;
  guarded(__func__, [&] {
    IndexData *ix = index_of(basis);
    if (ix == nullptr) return;
    BasisInfo const info = basis_info(basis);
    if (info.has_permutation_symmetries) ensure_norms(*ix, *info.group);
    if (representatives != nullptr) *representatives = ix->d_reps;
    if (norms != nullptr) *norms = ix->d_norms;
    if (count != nullptr) *count = (uint64_t)ix->number_states;
    status = 0;
  });
  return status;
}

}  // extern "C"
