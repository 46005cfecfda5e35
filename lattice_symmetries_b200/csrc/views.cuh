// views.cuh -- plain-old-data views of device-resident tables, passed to
// kernels by value, plus the scalar (one state per thread) device functions.
#pragma once

#include "common.cuh"

namespace lsb {

constexpr int kMaxDepth = 16;  // 2*log2(64)-1 = 11 stages for 64-bit states

// Symmetry group tables on the device (built from ls_hs_permutation_group,
// kernels/lattice_symmetries_types.h:109-119).
struct GroupView {
  int number_bits;
  int depth;
  int number_masks;    // |G|
  int spin_inversion;  // 0, +1, -1
  uint64_t flip_mask;
  uint64_t const *masks;  // [depth][|G|]
  double const *re;       // [|G|]
  double const *im;       // [|G|]
  uint8_t const *perm;    // [|G|][number_bits]: image bit i = source bit perm[i]
  // Distinct character values (value 0 is exactly 1+0i) and, per element, the
  // index of chi_j (low byte) and of spin_inversion * chi_j (high byte).
  double2 const *cvals;   // [number_chars]
  uint16_t const *cinfo;  // [|G|]
  int number_chars;       // 0 when there are more than 256 distinct values
  unsigned shifts[kMaxDepth];
};

// Sorted representatives + prefix bucket table (replaces
// ls_hs_state_index_binary_search_data, kernels/indexing.c:10-18).  Within a
// bucket all representatives share their top prefix bits, so the search only
// compares the low `shift` bits, kept in a compact side array (2 or 4 bytes per
// state instead of 8): the whole index of a 3e7-state basis is ~80 MB and
// stays L2-resident while x is gathered from HBM.
struct IndexView {
  uint64_t const *reps;
  int64_t number_states;
  uint32_t const *offsets32;  // [2^prefix + 1] when number_states < 2^32
  int64_t const *offsets64;   // otherwise
  uint16_t const *lows16;     // [number_states] low bits when shift <= 16
  uint32_t const *lows32;     // [number_states] low bits when 16 < shift <= 32
  uint64_t low_mask;          // (1 << shift) - 1
  int shift;                  // number_bits - prefix_bits
  int identity;               // state_index_is_identity: index == state
  uint64_t number_buckets;    // 2^prefix
  int steps;                  // bit_length(largest final window): trip count of the branchless search
  // Second level for crowded buckets (fixed-Hamming-weight representatives pile up
  // behind long runs of leading zeros: a few prefix buckets hold thousands of states
  // while the average holds eight).  sub_info[p] == 0: bucket p is searched directly;
  // otherwise (p2 << 27 | t): the next p2 bits below the prefix select entry k2 of the
  // table at subtab + 8 t, whose entries k2, k2 + 1 bracket the window relative to the
  // bucket's start.
  uint32_t const *sub_info;   // [2^prefix] or nullptr
  // subtab entries: (position relative to the bucket's start) << 6 | min(window length, 63); a length
  // field of 63 means "look at the next entry".  One 4-byte load brackets the window.
  uint32_t const *subtab;
  // level 1 as the hot kernels read it: {offsets32[p], sub_info[p] != 0 ? sub_info[p] : bucket length (< 2^27)}
  // -- one 8-byte load instead of three 4-byte ones (nullptr when there is no second level)
  uint2 const *entry8;
};

// Operator terms, structure-of-arrays on the device
// (ls_hs_nonbranching_terms, kernels/lattice_symmetries_types.h:140-151).
struct TermsView {
  int number_terms;
  double2 const *v;
  uint64_t const *m;
  uint64_t const *l;
  uint64_t const *r;
  uint64_t const *x;
  uint64_t const *s;
};

#if defined(__CUDACC__)

// Norm sums of valid states are exact small integers >= 1 (every stabiliser
// character is exactly 1.0); mathematically-zero sums in complex sectors come
// out as +-1e-16 noise.  Anything below 0.5 is treated as exactly zero.
constexpr double kNormThreshold = 0.5;

// Search window of `needle`: [lo, lo + n) is its prefix bucket (empty when the
// needle cannot be in the basis or `live` is false).
__device__ __forceinline__ void index_window(IndexView const &ix, uint64_t needle, bool live, int64_t &lo, int64_t &n) {
  lo = 0;
  n = 0;
  if (!live || ix.identity) return;
  uint64_t const p = needle >> ix.shift;
  if (ix.offsets32 != nullptr) {
    if (p < ix.number_buckets) {
      lo = (int64_t)__ldg(ix.offsets32 + p);
      n = (int64_t)__ldg(ix.offsets32 + p + 1) - lo;
    }
  } else if (ix.offsets64 != nullptr) {
    if (p < ix.number_buckets) {
      lo = __ldg(ix.offsets64 + p);
      n = __ldg(ix.offsets64 + p + 1) - lo;
    }
  } else {
    n = ix.number_states;
  }
  if (ix.sub_info != nullptr && p < ix.number_buckets) {
    uint32_t const s = __ldg(ix.sub_info + p);
    if (s != 0) {
      int const p2 = (int)(s >> 27);
      uint64_t const k2 = (needle >> (ix.shift - p2)) & ((uint64_t(1) << p2) - 1);
      uint32_t const *t = ix.subtab + (size_t)(s & 0x7ffffffu) * 8 + k2;
      uint32_t const v = __ldg(t);
      uint32_t cnt = v & 63u;
      if (cnt == 63u) cnt = (__ldg(t + 1) >> 6) - (v >> 6);
      lo += (int64_t)(v >> 6);
      n = (int64_t)cnt;
    }
  }
}

__device__ __forceinline__ uint64_t index_key_at(IndexView const &ix, int64_t i) {
  if (ix.lows16 != nullptr) return (uint64_t)__ldg(ix.lows16 + i);
  if (ix.lows32 != nullptr) return (uint64_t)__ldg(ix.lows32 + i);
  return __ldg(ix.reps + i);
}

// B independent lookups in lockstep (their loads overlap): index of needle[u]
// among the sorted representatives, or -1 (kernels/indexing.c:196-215, 273-325:
// prefix bucket, then a fixed-length branchless lower bound).
template <int B>
__device__ __forceinline__ void index_find(IndexView const &ix, uint64_t const (&needle)[B], bool const (&live)[B],
                                           int64_t (&found)[B]) {
  if (ix.identity) {
#pragma unroll
    for (int u = 0; u < B; ++u) found[u] = live[u] ? (int64_t)needle[u] : (int64_t)-1;
    return;
  }
  bool const compact = ix.lows16 != nullptr || ix.lows32 != nullptr;
  int64_t lo[B], n[B], end[B];
  uint64_t key[B];
#pragma unroll
  for (int u = 0; u < B; ++u) {
    index_window(ix, needle[u], live[u], lo[u], n[u]);
    end[u] = lo[u] + n[u];
    key[u] = compact ? (needle[u] & ix.low_mask) : needle[u];
  }
#pragma unroll 1
  for (int s = 0; s < ix.steps; ++s) {
#pragma unroll
    for (int u = 0; u < B; ++u) {
      int64_t const half = n[u] >> 1;
      int64_t const mid = lo[u] + half;
      bool less = false;
      if (n[u] > 0) less = index_key_at(ix, mid) < key[u];
      lo[u] = less ? mid + 1 : lo[u];
      n[u] = less ? n[u] - half - 1 : half;
    }
  }
#pragma unroll
  for (int u = 0; u < B; ++u) {
    found[u] = -1;
    // lo[u] is the lower bound inside the needle's bucket; it may sit at the bucket's end
    if (live[u] && lo[u] < end[u] && index_key_at(ix, lo[u]) == key[u]) found[u] = lo[u];
  }
}

// Lean variant for the hot kernels: 32-bit positions (offsets32 tables), keys of a
// fixed type Low (uint16_t / uint32_t low bits, or uint64_t full states).  Upper-bound
// search by descending powers of two that carries the last key <= needle along, so the
// membership test needs no extra load; `steps` probes, each a predicated load.
template <class Low> __device__ __forceinline__ Low const *index_keys(IndexView const &ix);
template <> __device__ __forceinline__ uint16_t const *index_keys<uint16_t>(IndexView const &ix) { return ix.lows16; }
template <> __device__ __forceinline__ uint32_t const *index_keys<uint32_t>(IndexView const &ix) { return ix.lows32; }
template <> __device__ __forceinline__ uint64_t const *index_keys<uint64_t>(IndexView const &ix) { return ix.reps; }

// Wide = true: the index of a basis with 2^32 or more states (level 1 = offsets64, no second level): window
// positions stay 32-bit, relative to the 64-bit start of the needle's bucket.
template <class Low, int B, bool Wide = false>
__device__ __forceinline__ void index_find32(IndexView const &ix, uint64_t const (&needle)[B], bool const (&live)[B],
                                             int64_t (&found)[B]) {
  Low const *__restrict__ keys = index_keys<Low>(ix);
  uint32_t lo[B], pos[B], end[B];
  int64_t base[B];
  Low key[B], val[B];
#pragma unroll
  for (int u = 0; u < B; ++u) {
    uint64_t const p = needle[u] >> ix.shift;
    uint32_t l = 0, n = 0;
    base[u] = 0;
    if (live[u] && p < ix.number_buckets) {
      if constexpr (Wide) {
        int64_t const b0 = __ldg(ix.offsets64 + p);
        base[u] = b0;
        n = (uint32_t)(__ldg(ix.offsets64 + p + 1) - b0);
      } else {
        uint32_t s = 0;
        if (ix.entry8 != nullptr) {
          uint2 const e = __ldg(ix.entry8 + p);
          l = e.x;
          n = e.y;
          s = (e.y >> 27) != 0 ? e.y : 0u;  // p2 >= 1 sits in the top five bits; plain lengths are < 2^27
        } else {
          l = __ldg(ix.offsets32 + p);
          n = __ldg(ix.offsets32 + p + 1) - l;
          s = ix.sub_info != nullptr ? __ldg(ix.sub_info + p) : 0u;
        }
        if (s != 0) {
          int const p2 = (int)(s >> 27);
          uint32_t const k2 = (uint32_t)(needle[u] >> (ix.shift - p2)) & ((1u << p2) - 1u);
          uint32_t const *t = ix.subtab + (size_t)(s & 0x7ffffffu) * 8 + k2;
          uint32_t const v = __ldg(t);
          n = v & 63u;
          if (n == 63u) n = (__ldg(t + 1) >> 6) - (v >> 6);
          l += v >> 6;
        }
      }
    }
    lo[u] = pos[u] = l;
    end[u] = l + n;
    key[u] = sizeof(Low) < 8 ? (Low)(needle[u] & ix.low_mask) : (Low)needle[u];
    val[u] = 0;
  }
#pragma unroll 1
  for (uint32_t s = ix.steps > 0 ? (1u << (ix.steps - 1)) : 0u; s != 0; s >>= 1) {
#pragma unroll
    for (int u = 0; u < B; ++u) {
      uint32_t const cand = pos[u] + s;
      if (cand <= end[u]) {
        Low const v = Wide ? __ldg(keys + base[u] + (cand - 1)) : __ldg(keys + (cand - 1));
        if (v <= key[u]) {
          pos[u] = cand;
          val[u] = v;
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < B; ++u)
    found[u] = (pos[u] > lo[u] && val[u] == key[u]) ? (Wide ? base[u] + (int64_t)(pos[u] - 1) : (int64_t)(pos[u] - 1))
                                                    : (int64_t)-1;
}

__device__ __forceinline__ int64_t state_index(IndexView const &ix, uint64_t needle) {
  uint64_t const needles[1] = {needle};
  bool const live[1] = {true};
  int64_t found[1];
  index_find<1>(ix, needles, live, found);
  return found[0];
}

// Scalar orbit walk: apply every group element's Benes network to x
// (kernels/generator.cpp:5-8, :88-94), masks staged in shared memory as W
// (uint32_t when number_bits <= 32, halving the integer work).
template <class W>
struct OrbitScalar {
  GroupView const &g;
  W const *smasks;  // shared: [depth][|G|]
  __device__ __forceinline__ W image(W x, int j) const {
    W y = x;
#pragma unroll 1
    for (int k = 0; k < g.depth; ++k) y = bit_permute_step<W>(y, smasks[k * g.number_masks + j], g.shifts[k]);
    return y;
  }
};

template <class W>
__device__ __forceinline__ void stage_masks(GroupView const &g, W *smasks) {
  int const total = g.depth * g.number_masks;
  for (int i = threadIdx.x; i < total; i += blockDim.x) smasks[i] = (W)g.masks[i];
  __syncthreads();
}

// state_info of one state (kernels/generator.cpp:24-54, 77-140): representative,
// character of the element that maps x onto it, stabiliser character sum n.
template <class W>
__device__ __forceinline__ void state_info_scalar(GroupView const &g, W const *smasks, uint64_t x64,
                                                  uint64_t &rep, double &c_re, double &c_im,
                                                  double &n) {
  OrbitScalar<W> orbit{g, smasks};
  W const x = (W)x64;
  W const flip = (W)g.flip_mask;
  W r = x;
  int best = -1;  // -1: identity/no change; j: element j; j + |G|: element j with flip
  double acc = 0.0;
  int const G = g.number_masks;
  int const inv = g.spin_inversion;
#pragma unroll 1
  for (int j = 0; j < G; ++j) {
    W const y = orbit.image(x, j);
    if (y < r) { r = y; best = j; }
    if (y == x) acc += g.re[j];
    if (inv != 0) {
      W const yf = y ^ flip;
      if (yf < r) { r = yf; best = j + G; }
      if (yf == x) acc += (double)inv * g.re[j];
    }
  }
  rep = (uint64_t)r;
  if (best < 0) { c_re = 1.0; c_im = 0.0; }
  else if (best < G) { c_re = g.re[best]; c_im = g.im[best]; }
  else { c_re = (double)inv * g.re[best - G]; c_im = (double)inv * g.im[best - G]; }
  n = acc;
}

__device__ __forceinline__ double norm_from_sum(GroupView const &g, double n) {
  // kernels/generator.cpp:135-140, with the noise band clamped to zero.
  if (g.number_masks > 0 && n < kNormThreshold) return 0.0;
  return sqrt(n / (double)((g.spin_inversion == 0 ? 1 : 2) * g.number_masks));
}

#endif  // __CUDACC__

}  // namespace lsb
