// state.hpp -- host-side objects that own the device-resident tables.
#pragma once

#include <memory>
#include <unordered_map>
#include <vector>

#include "views.cuh"

namespace lsb {

constexpr uint32_t kGroupMagic = 0x4c534731u;  // "LSG1"
constexpr uint32_t kIndexMagic = 0x4c534931u;  // "LSI1"

// The object behind ls_hs_basis_kernels::{state_info_data,is_representative_data};
// replaces ls_internal_halide_kernel_data (kernels/kernels.c:100-110).  Unlike
// the reference (which borrows Haskell-owned arrays, kernels.c:121-156) the
// tables are copied: to the host vectors below and to the device.
struct GroupData {
  uint32_t magic = kGroupMagic;
  int number_bits = 0;
  int depth = 0;
  int number_masks = 0;
  int spin_inversion = 0;
  uint64_t flip_mask = 0;
  uint64_t id = 0;  // unique per object; tags the constant-memory cache
  bool real_characters = true;
  std::vector<uint64_t> masks, shifts;
  std::vector<double> re, im;
  std::vector<uint8_t> perm;  // [|G|][number_bits], recovered from the networks
  std::vector<double> cvals;      // distinct character values, interleaved (re, im); cvals[0] = 1+0i
  std::vector<uint16_t> cinfo;    // per element: index of chi_j | index of (inversion * chi_j) << 8
  double2 *d_cvals = nullptr;
  uint16_t *d_cinfo = nullptr;
  uint64_t *d_masks = nullptr;
  double *d_re = nullptr, *d_im = nullptr;
  uint8_t *d_perm = nullptr;

  GroupView view() const;
  ~GroupData();
};

// The object behind ls_hs_basis_kernels::state_index_data; replaces
// ls_hs_state_index_binary_search_data (kernels/indexing.c:10-18).
struct IndexData {
  uint32_t magic = kIndexMagic;
  int64_t number_states = 0;
  uint64_t const *host_reps = nullptr;  // borrowed, may be NULL (device-only basis)
  int number_bits = 0;
  int prefix_bits = 0;
  int shift = 0;
  int steps = 0;  // bit_length(largest bucket)
  bool identity = false;
  uint64_t *d_reps = nullptr;
  bool owns_d_reps = true;
  uint32_t *d_offsets32 = nullptr;
  int64_t *d_offsets64 = nullptr;
  uint16_t *d_lows16 = nullptr;  // low `shift` bits of every representative (shift <= 16)
  uint32_t *d_lows32 = nullptr;  // (16 < shift <= 32)
  uint32_t *d_sub_info = nullptr;  // second-level tables of the crowded buckets (IndexView)
  uint32_t *d_subtab = nullptr;
  uint2 *d_entry8 = nullptr;       // level 1 packed for the hot kernels (IndexView::entry8)
  double *d_norms = nullptr;  // state_info norms of the representatives (lazy)

  IndexView view() const;
  ~IndexData();
};

// Device copies adopted from a basis build, keyed by the host pointer handed
// out in chpl_external_array::elts, until an IndexData claims them.
struct BuiltReps {
  uint64_t *d_reps = nullptr;
  double *d_norms = nullptr;
  uint64_t count = 0;
};
std::unordered_map<void const *, BuiltReps> &built_registry();

// Device copy of one ls_hs_nonbranching_terms table
// (kernels/lattice_symmetries_types.h:140-151), structure of arrays.
struct TermsDev {
  int number_terms = 0;
  std::vector<double> v;  // host copy, interleaved (re, im); used to detect changes
  std::vector<uint64_t> m, l, r, x, s;
  double2 *d_v = nullptr;
  uint64_t *d_m = nullptr, *d_l = nullptr, *d_r = nullptr, *d_x = nullptr, *d_s = nullptr;
  bool same_as(ls_hs_nonbranching_terms const *t) const;
  void upload(ls_hs_nonbranching_terms const *t);
  void release();
  TermsView view() const;
  ~TermsDev() { release(); }
};

// Per-operator device state, cached by operator address and revalidated
// against the caller's term tables on every use.
struct OperatorDev {
  TermsDev off, diag;
  int distinct_x = 0;  // Operator.hs:190-193 maxNumberOffDiag
  uint64_t version = 1;  // bumped whenever the term tables are re-uploaded
  // matrix-element statistics of the basis the operator was last used with
  void const *stats_index = nullptr;
  int64_t stats_rows = -1;
  int64_t stats_elements = 0;
};
OperatorDev &operator_dev(ls_hs_operator const *op);

// Basis predicates the Chapel side asks the Haskell host for
// (haskell/src/LatticeSymmetries/Basis.hs:701-774), derived from the struct.
struct BasisInfo {
  int number_bits;
  bool fixed_hamming;
  int hamming_weight;       // spin: number_up; fermions: number_particles
  bool has_permutation_symmetries;
  bool has_spin_inversion;
  bool spinful_sectors;     // spinful fermions with (n_up, n_down)
  uint64_t min_state, max_state;
  GroupData const *group;   // may be NULL
};
BasisInfo basis_info(ls_hs_basis const *basis);
IndexData *index_of(ls_hs_basis const *basis);
// An index object is going away: cached per-index state of the matvec (phased canonicalisation) must not outlive it.
void matvec_forget_index(IndexData const *ix);

// Combinadics (haskell/src/LatticeSymmetries/Basis.hs:487-550)
uint64_t binomial(int n, int k);
uint64_t fixed_hamming_state_to_index(uint64_t state);
uint64_t fixed_hamming_index_to_state(uint64_t index, int hamming_weight);

}  // namespace lsb
