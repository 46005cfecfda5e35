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
// ls_hs_state_index_binary_search_data, kernels/indexing.c:10-18).
struct IndexView {
  uint64_t const *reps;
  int64_t number_states;
  uint32_t const *offsets32;  // [2^prefix + 1] when number_states < 2^32
  int64_t const *offsets64;   // otherwise
  int shift;                  // number_bits - prefix_bits
  int identity;               // state_index_is_identity: index == state
  uint64_t number_buckets;    // 2^prefix
  int steps;                  // bit_length(largest bucket): trip count of the branchless search
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

// lower_bound-style bucketed search; returns index or -1
// (kernels/indexing.c:196-215, :273-325).
__device__ __forceinline__ int64_t state_index(IndexView const &ix, uint64_t needle) {
  if (ix.identity) return (int64_t)needle;
  uint64_t const p = needle >> ix.shift;
  int64_t lo, hi;
  if (ix.offsets32 != nullptr) {
    if (p >= ix.number_buckets) return -1;
    lo = (int64_t)__ldg(ix.offsets32 + p);
    hi = (int64_t)__ldg(ix.offsets32 + p + 1);
  } else if (ix.offsets64 != nullptr) {
    if (p >= ix.number_buckets) return -1;
    lo = __ldg(ix.offsets64 + p);
    hi = __ldg(ix.offsets64 + p + 1);
  } else {
    lo = 0;
    hi = ix.number_states;
  }
  while (lo < hi) {
    int64_t const mid = (lo + hi) >> 1;
    uint64_t const v = __ldg(ix.reps + mid);
    if (v < needle) lo = mid + 1; else hi = mid;
  }
  return (lo < ix.number_states && __ldg(ix.reps + lo) == needle) ? lo : (int64_t)-1;
}

// Search window of `needle`: [lo, lo + n) is its prefix bucket (empty when the
// needle cannot be in the basis or `live` is false).  Followed by ix.steps
// rounds of  half = n >> 1; mid = lo + half; reps[mid] < needle ? (lo = mid + 1,
// n -= half + 1) : (n = half)  -- the fixed-length branchless lower bound of
// kernels/indexing.c:196-215 -- after which reps[lo] == needle decides.
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
