// plane_table.cuh -- the per-element "plane renaming" table of the bit-sliced
// kernels, held in constant memory.
//
// For group element j and output bit i the image of a state takes its bit i
// from source bit perm_j[i] (Benes.hs:338-345 permuteBits').  With 32 states
// transposed into bit planes stored column-wise in shared memory, plane i of
// the image is simply plane perm_j[i] of the input, i.e. a byte offset
// perm_j[i] * (threads per block) * 4 from the thread's own column.  The table
// index is warp-uniform, so the offsets arrive through the uniform datapath
// (LDCU) and cost no ALU/LSU issue slots.
//
// Constant memory is per translation unit: every .cu that includes this header
// owns a private copy and uploads it on demand.
#pragma once

#include <algorithm>
#include <vector>

#include "state.hpp"

namespace lsb {

constexpr int kMaxPermTable = 15360;  // uint32 entries (60 KB of constant memory), read as uint4 (LDCU.128)
static __constant__ uint4 c_plane_rows[kMaxPermTable / 4];
static uint64_t g_plane_table_owner = 0;  // GroupData::id, NP and block size currently resident

// Row j of the table (kPlaneRowExtra + np 32-bit entries, 16-byte aligned):
//   [0, np)   byte offset of the source plane of output plane i
//   [np]      byte offset of the source plane of the TOP live bit (number_bits - 1):
//             the spin-flipped image ~y is smaller than y iff that bit of y is set
//   [np + 1]  character indices of element j (GroupData::cinfo)
//   [np + 2]  same as [np] when the flip plane differs from the previous row's, else kNoRetarget
//   [np + 3]  padding
// Planes are padded to a multiple of four (np >= number_bits); padding planes
// map onto themselves.  Returns false when the table does not fit.
constexpr int kPlaneRowExtra = 4;
constexpr uint32_t kNoRetarget = 0xffffffffu;
static inline bool upload_plane_offsets(GroupData const &g, int np, int threads_per_block) {
  size_t const stride = (size_t)np + kPlaneRowExtra;
  size_t const entries = (size_t)g.number_masks * stride;
  if (entries > (size_t)kMaxPermTable) return false;
  if ((np & 3) != 0) return false;
  uint64_t const tag = (g.id << 20) | ((uint64_t)np << 12) | (uint64_t)threads_per_block;
  if (g_plane_table_owner == tag) return true;
  std::vector<uint32_t> table(entries, 0u);
  // Rows are ordered by the source of the top bit (the identity's class first, so that an
  // identity element stays in row 0): consecutive rows then share their flip plane and the
  // kernels re-target the flipped copy of the planes only when it changes.  The orbit
  // minimum / stabiliser sums do not depend on the order of the elements.
  auto top_of = [&](int j) { return g.number_bits > 0 ? (int)g.perm[(size_t)j * g.number_bits + (g.number_bits - 1)] : 0; };
  std::vector<int> order((size_t)g.number_masks);
  for (int j = 0; j < g.number_masks; ++j) order[(size_t)j] = j;
  int const first_top = g.number_masks > 0 ? top_of(0) : 0;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    int const ka = top_of(a) == first_top ? -1 : top_of(a), kb = top_of(b) == first_top ? -1 : top_of(b);
    return ka < kb;
  });
  int previous_top = -1;
  for (int row = 0; row < g.number_masks; ++row) {
    int const j = order[(size_t)row];
    for (int i = 0; i < np; ++i) {
      int const src = i < g.number_bits ? g.perm[(size_t)j * g.number_bits + i] : i;
      table[(size_t)row * stride + i] = (uint32_t)(src * threads_per_block * 4);
    }
    int const top = top_of(j);
    table[(size_t)row * stride + np] = (uint32_t)(top * threads_per_block * 4);
    table[(size_t)row * stride + np + 1] = g.cinfo.empty() ? 0u : (uint32_t)g.cinfo[(size_t)j];
    table[(size_t)row * stride + np + 2] = top != previous_top ? (uint32_t)(top * threads_per_block * 4) : kNoRetarget;
    previous_top = top;
  }
  CUDA_CHECK(cudaMemcpyToSymbolAsync(c_plane_rows, table.data(), entries * sizeof(uint32_t), 0,
                                     cudaMemcpyHostToDevice, runtime().stream));
  CUDA_CHECK(cudaStreamSynchronize(runtime().stream));  // table is a stack temporary
  g_plane_table_owner = tag;
  return true;
}

#if defined(__CUDACC__)
// Row j of the table in (uniform) registers: NP + kPlaneRowExtra offsets, fetched
// four at a time.  j must be warp-uniform.
template <int NP>
struct PlaneRow {
  uint32_t off[NP + kPlaneRowExtra];
  __device__ __forceinline__ explicit PlaneRow(int j) {
    uint4 const *row = c_plane_rows + (size_t)j * ((NP + kPlaneRowExtra) / 4);
#pragma unroll
    for (int q = 0; q < (NP + kPlaneRowExtra) / 4; ++q) {
      uint4 const v = row[q];
      off[4 * q] = v.x;
      off[4 * q + 1] = v.y;
      off[4 * q + 2] = v.z;
      off[4 * q + 3] = v.w;
    }
  }
  __device__ __forceinline__ uint32_t operator[](int i) const { return off[i]; }
};
#endif

// True when element 0 is the identity with character exactly 1 (always the
// case for groups built by Group.hs:174-183, whose ascending order puts the
// identity permutation first); lets kernels skip it.
static inline bool identity_is_first(GroupData const &g) {
  if (g.number_masks == 0 || g.re[0] != 1.0 || g.im[0] != 0.0) return false;
  for (int i = 0; i < g.number_bits; ++i)
    if (g.perm[(size_t)i] != i) return false;
  return true;
}

}  // namespace lsb
