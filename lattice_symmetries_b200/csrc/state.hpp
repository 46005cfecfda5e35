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
  struct DistShard *dist = nullptr;  // set when these are the LOCAL rows of a basis sharded over ranks (dist.cu); owned

  IndexView view() const;
  ~IndexData();
};

// A basis sharded over the ranks of a communicator (dist.cu): rank r owns the contiguous range
// [bounds[r], bounds[r + 1]) of the globally sorted representatives -- its IndexData holds exactly those rows.
// Replaces the hash distribution of chapel/src/StatesEnumeration.chpl:198-212.
struct DistShard {
  int world = 1, rank = 0;
  int64_t dim = 0;                  // representatives over all ranks
  std::vector<int64_t> bounds;      // [world + 1]
  std::vector<uint64_t> splitters;  // [world] first representative of every rank (~0 for an empty rank)
  uint64_t *d_splitters = nullptr;
  // all-gather products: state -> GLOBAL row over the whole basis (compact keys + level 1 only, replicated on every
  // rank) and the replicated pre-scaled vector.  nullptr: not built (does not fit, or all-to-all was asked for).
  IndexData *global_index = nullptr;
  double *d_xs_full = nullptr;
  size_t xs_full_words = 0;
  // all-to-all products: records grouped by owner, their per-owner displacements, the received records
  struct PushBuffers *push = nullptr;
  ~DistShard();
};

// What the rows and the lookups of a product refer to (matvec_device).  nullptr = the basis' own list for both.
struct MvTarget {
  IndexView index;        // ranks a representative -> position in xs
  uint64_t const *rows;   // [number_rows] representatives of the rows of this call
  double const *norms;    // [number_rows], or nullptr when all 1
  int64_t number_rows;
  double const *xs;       // replicated n_j x_j (x_j when norms == nullptr), indexed by `index`
  cudaEvent_t xs_ready;   // nullptr, or: xs is complete once this event fires (the all-gather runs on another stream
                          // under the canonicalisation of the first chunk, which does not read it)
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
  // sorted ranking: chunks of rows cut by exact matrix-element counts (matvec_device), cached per row set
  struct ChunkPlan {
    void const *rows = nullptr;
    int64_t row_begin = -1, row_end = -1, capacity = 0;
    uint64_t version = 0;
    std::vector<int64_t> begin;     // [chunks + 1] first row of every chunk
    std::vector<int64_t> elements;  // [chunks] off-diagonal matrix elements of every chunk
  } plan;
};
OperatorDev &operator_dev(ls_hs_operator const *op);

// ---- entry points of matvec.cu used by the distributed drivers (dist.cu) --------------------------------------
void matvec_device(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, double const *d_x, double *d_y,
                   bool complex_vectors, int number_vectors = 1, int64_t x_stride = 0, int64_t y_stride = 0,
                   double *host_y = nullptr, int phase = 0, MvTarget const *target = nullptr);
bool matvec_finish();
bool matvec_last_row_costs(int64_t number_rows, int segments, double *out);
extern char const *kInvalidIndexMessage;
void launch_prescale(int64_t n, bool complex_vectors, double const *norms, double const *x, double *xs,
                     cudaStream_t stream = nullptr);
int64_t count_elements(OperatorDev &od, uint64_t const *d_rows, int64_t row_begin, int64_t row_end);
void release_count_scratch(size_t above_bytes);
void count_elements_segments(OperatorDev &od, uint64_t const *d_rows, int64_t number_rows, uint64_t const *d_starts,
                             int64_t number_segments, uint64_t *d_out);

// Push form (all-to-all products, mirrors chapel/src/DistributedMatrixVector.chpl:545-579, 775-807): records
// (representative, coefficient) grouped by the rank that owns the representative.
struct PushRecord {
  uint64_t rep;
  double c;
};
struct PushBuffers {
  DeviceBuffer<PushRecord> send, recv;
  DeviceBuffer<uint32_t> hist, base;
  DeviceBuffer<unsigned char> scan_tmp;
  DeviceBuffer<unsigned long long> displs;  // [world + 1] start of every owner's records in `send`
};
int64_t push_chunk_rows(ls_hs_operator const *op, int64_t local_rows);
void push_begin(ls_hs_operator const *op, IndexData const &local, double const *d_x, double *d_y);
void push_produce(ls_hs_operator const *op, IndexData const &local, DistShard &shard, int64_t chunk_begin,
                  int64_t chunk_rows, double const *d_x);
void push_consume(ls_hs_operator const *op, IndexData const &local, PushRecord const *records, int64_t count,
                  double *d_y);

// ---- entry points of basis_build.cu / index.cu used by dist.cu -------------------------------------------------
struct BuildResult {
  uint64_t *d_reps = nullptr;
  double *d_norms = nullptr;  // nullptr for unprojected bases
  uint64_t count = 0;
  uint64_t const *d_block_starts = nullptr;  // cyclic shares: [number_blocks + 1] first row of every block (library scratch,
                                             // valid until the next build)
};
// Allocation of the array that becomes basis->representatives: managed (device-resident, host-readable) when the
// host view is wanted and supported, plain device memory otherwise.
uint64_t *alloc_representatives(uint64_t count);
using Ranges = std::vector<std::pair<uint64_t, uint64_t>>;
// A rank's block-cyclic share of the candidate range: blocks of 32 << shift candidates, block b scanned by rank
// b % world (the last block may be partial).
struct CyclicShare {
  int shift = 15, world = 1, rank = 0;
  uint64_t blocks_total = 0;       // over all ranks
  uint64_t number_blocks = 0;      // of this rank
  uint64_t virtual_candidates = 0; // candidates in this rank's blocks
};
CyclicShare cyclic_share(uint64_t total, int block_shift, int world, int rank);
BuildResult build_ranges(ls_hs_basis const *basis, Ranges ranges, std::vector<uint64_t> *counts = nullptr,
                         bool host_visible = false, CyclicShare const *cyc = nullptr);
uint64_t number_candidates(ls_hs_basis const *basis);
// Installs device-resident representatives (+ norms) as the basis' list: host view, index, kernels.
void install_representatives(ls_hs_basis *basis, uint64_t *d_reps, double *d_norms, uint64_t count, int cache_bits);
int index_choose_prefix_bits(int64_t n, int number_bits);
void index_set_lean(bool lean);
void index_local_keys(uint64_t const *d_reps, int64_t n, int shift, void *d_keys, int key_bytes);
void index_local_offsets64(uint64_t const *d_reps, int64_t n, int shift, int64_t number_offsets, int64_t *d_out);
int index_steps_from_offsets64(int64_t const *d_offsets, int64_t number_buckets);
IndexData *create_index_from_device(uint64_t *d_reps, int64_t count, int number_bits, int prefix_bits);

// ---- entry points of dist.cu ------------------------------------------------------------------------------------------
int comm_world();  // ranks of the active communicator (1 without one)
bool dist_is_sharded(ls_hs_basis const *basis);
// y_local = (H x)_local on this rank's rows; x, y device-resident, local length.  mode: 0 auto, 1 all-gather, 2 all-to-all.
// x_ready (optional): d_x is complete once this event fires (an upload in flight on another stream); host_y (optional,
// pinned): finished row chunks of d_y are also copied there on the copy stream.
void dist_matvec_local(ls_hs_operator const *op, double const *d_x, double *d_y, int mode, bool complex_vectors,
                       cudaEvent_t x_ready = nullptr, double *host_y = nullptr);
void dist_build_local(ls_hs_basis *basis, ls_hs_operator const *balance_for, int flags);

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
