// matvec_args.cuh -- argument block, staged term tables and the rare-path device helpers shared by the
// matvec translation units (matvec.cu: drivers + memory-bound kernels; orbit.cu: the integer-bound
// canonicalisation kernels, compiled separately because their 64 template instances dominate build time).
#pragma once

#include <type_traits>

#include "bitslice.cuh"
#include "state.hpp"

namespace lsb {

#ifndef LS_ORBIT_RETARGET
#define LS_ORBIT_RETARGET 1  // orbit_kernel: re-target the flipped planes in place instead of one XOR per plane
#endif
constexpr int kOrbitThreads = 128;   // one thread = one word of 32 matrix elements
constexpr int kGatherThreads = 128;  // one thread = one row
constexpr int kMvIdxPlanes = 8;      // bit-sliced path: at most 256 distinct character values
constexpr int kGatherBatch = 4;      // independent searches in flight per thread

enum : int { kModeNone = 0, kModeInversion = 1, kModeGroup = 2, kModeGroupScalar = 3 };

// Everything one chunk of rows needs; passed by value.
struct MatvecArgs {
  GroupView g;
  IndexView ix;
  uint64_t const *rows;  // representatives of the rows, rows[chunk_begin + r] (a distributed product reads its local
                         // shard here while ix ranks against the whole basis; otherwise == ix.reps)
  TermsView off, diag;
  int mode;
  int number_idx_planes;  // ceil(log2(number of distinct characters))
  int debug_skip;         // LS_B200_MV_SKIP (profiling only): 1 = no orbit walk, 2 = no search/gather
  int number_chars;
  double2 const *cvals;  // distinct character values; chars[cidx] in the kernels
  int complex_vectors;   // x, xs, y hold interleaved (re, im)
  int spin_inversion;
  uint64_t inversion_mask;
  int64_t row_begin;     // first row of the CALL (y[0] is this row)
  int64_t chunk_begin;   // first row of this chunk
  int chunk_rows;
  double const *norms;   // n_i of the representatives; nullptr when all 1
  double const *x;       // caller's vector (diagonal part)
  double const *xs;      // n_j x[j] (== x when norms is nullptr)
  double *y;
  int *error_flag;
  // chunk intermediates
  uint32_t *counts;      // [chunk_rows + 1] matches per row (last = 0)
  uint32_t *offsets;     // [chunk_rows + 1] exclusive scan of counts
  uint64_t *q_rep;       // [capacity] representative of every matrix element, CSR order
  uint8_t *q_cidx;       // [capacity] index of the minimising character
  uint16_t *q_tsign;     // [capacity] split path: term | sign << 15 of every matrix element (nullptr: not recorded)
  double *vals;          // [capacity] (x2 when complex) fused path: conj(chi) w sign n_j x_j of every matrix element
  // sorted ranking (large bases, see matvec_device): q_sorted[k] = the k-th smallest representative of the chunk (by its
  // leading bits), perm[k] = its position in CSR order; nullptr: rank in CSR order
  uint64_t const *q_sorted;
  uint32_t const *perm;
  // block matvec (split path): vector v reads x + v x_stride / xs + v x_stride, writes y + v y_stride and
  // vals + v vals_stride (strides in scalars of the vector type)
  int number_vectors;
  int64_t x_stride, xs_stride, y_stride, vals_stride;
};

// Stabiliser character sum of x read straight from the global tables; used on
// the (rare) path that decides whether a missing index is an error.
static __device__ __noinline__ double stabiliser_sum_global(GroupView g, uint64_t x) {
  double acc = 0.0;
  for (int j = 0; j < g.number_masks; ++j) {
    uint64_t y = x;
    for (int k = 0; k < g.depth; ++k)
      y = bit_permute_step<uint64_t>(y, __ldg(g.masks + (size_t)k * g.number_masks + j), g.shifts[k]);
    if (y == x) acc += __ldg(g.re + j);
    if (g.spin_inversion != 0 && (y ^ g.flip_mask) == x) acc += (double)g.spin_inversion * __ldg(g.re + j);
  }
  return acc;
}

// Scalar orbit minimum from the global tables (kModeGroupScalar: groups that
// do not fit the bit-sliced path, and A/B validation via LS_B200_MATVEC=scalar).
static __device__ __noinline__ void orbit_min_global(GroupView g, uint64_t x, uint64_t &rep, int &element, int &flipped) {
  uint64_t r = x;
  int best = -1, fl = 0;
  for (int j = 0; j < g.number_masks; ++j) {
    uint64_t y = x;
    for (int k = 0; k < g.depth; ++k)
      y = bit_permute_step<uint64_t>(y, __ldg(g.masks + (size_t)k * g.number_masks + j), g.shifts[k]);
    if (y < r) { r = y; best = j; fl = 0; }
    if (g.spin_inversion != 0) {
      uint64_t const yf = y ^ g.flip_mask;
      if (yf < r) { r = yf; best = j; fl = 1; }
    }
  }
  rep = r;
  element = best;
  flipped = fl;
}

// Shared-memory copy of the adjoint off-diagonal terms: match on l, weight
// w = v (-1)^{|x&s|}.
struct AdjointTerms {
  uint64_t *m, *l, *x, *s;
  double2 *w;
  int T;
  static __host__ __device__ size_t bytes(int T, bool with_weights) { return (size_t)T * (with_weights ? 48 : 24); }
  __device__ void stage(unsigned char *base, TermsView const &off, bool with_weights) {
    T = off.number_terms;
    m = reinterpret_cast<uint64_t *>(base);
    l = m + T;
    x = l + T;
    s = x + T;
    w = reinterpret_cast<double2 *>(s + T);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      uint64_t const xt = off.x[t];
      m[t] = off.m[t];
      l[t] = off.l[t];
      x[t] = xt;
      if (with_weights) {
        uint64_t const st = off.s[t];
        double2 v = off.v[t];
        if (__popcll(xt & st) & 1) { v.x = -v.x; v.y = -v.y; }
        s[t] = st;
        w[t] = v;
      }
    }
  }
};

// shared-memory geometry of the orbit kernels (the drivers size the launches)
constexpr int kWarpSlabBytes = 32 * 32 * 8;
#ifndef LS_FUSED_BATCH
#define LS_FUSED_BATCH 8
#endif
constexpr int kFusedBatch = LS_FUSED_BATCH;  // independent searches in flight per lane
constexpr int kOutPitch = 33;                          // u64 words per row of the [k][owner] output slab
constexpr int kCidxPitch = 36;                         // bytes per row of the [k][owner] character-index slab
constexpr int kFusedSlabBytes = 32 * kOutPitch * 8;    // >= kWarpSlabBytes
constexpr int kFusedTsignBytes = 1024 * 2;             // term | sign << 15 per element
constexpr int kFusedCidxBytes = 32 * kCidxPitch;
constexpr int kFusedWarpBytes = kFusedSlabBytes + kFusedTsignBytes + kFusedCidxBytes;
static_assert(kFusedSlabBytes >= kWarpSlabBytes && kFusedWarpBytes % 16 == 0, "slab layout");

// orbit.cu
// Uploads the plane table of g (constant memory of orbit.cu); false when the bit-sliced path cannot be used.
bool orbit_prepare(GroupData const &g, int np);
// Launches orbit_kernel<NP, INV> (fused: orbit_gather_kernel) over `words` 32-element words of the chunk.
void orbit_launch(int np, bool inv, bool fused, size_t words, size_t smem, cudaStream_t stream, MatvecArgs const &a);

}  // namespace lsb
