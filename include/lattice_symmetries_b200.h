/*
 * lattice_symmetries_b200.h -- C ABI of the B200-native hot-path library.
 *
 * liblattice_symmetries_b200.so is a drop-in for the two native pieces the
 * reference's Haskell host and Python cffi wrapper link against on this path:
 *
 *   (1) libkernels.a            kernels/{kernels.c,indexing.c,reference.c}
 *   (2) liblattice_symmetries_chapel.so   chapel/src/ (all .chpl)
 *
 * Every entry point below names the reference declaration it replaces
 * (file:line in twesterhout/lattice-symmetries @ 36215fe).  Struct layouts are
 * field-for-field those of kernels/lattice_symmetries_types.h (the contract
 * the Haskell Storable instances in haskell/src/LatticeSymmetries/FFI.hs:69-143
 * marshal into); static_asserts in csrc/abi_layout.cu pin the offsets.
 *
 * All pointers in the ls_hs_ / ls_chpl_ / ls_internal_ entry points are HOST
 * pointers, exactly as in the reference.  The ls_b200_ entry points at the
 * bottom are extensions for callers that keep vectors resident in HBM.
 *
 * There is NO CPU fallback: every compute entry point runs CUDA kernels for
 * sm_100a and reports failure through ls_hs_error() when no device is usable.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types: kernels/lattice_symmetries_types.h ------------------------- */

/* :27-32 (Chapel's chpl-external-array.h when present) */
typedef struct chpl_external_array {
  void *elts;
  uint64_t num_elts;
  void *freer; /* void (*)(void*) called with elts, or NULL when borrowed */
} chpl_external_array;

/* :43-56 -- interleaved (re, im); layout-compatible with _Complex double and
 * std::complex<double>. */
typedef struct ls_hs_scalar {
  double _real;
  double _imag;
} ls_hs_scalar;

/* :81-85 */
typedef enum ls_hs_particle_type {
  LS_HS_SPIN,
  LS_HS_SPINFUL_FERMION,
  LS_HS_SPINLESS_FERMION
} ls_hs_particle_type;

/* :87-98 */
typedef void (*ls_hs_internal_state_index_kernel_type)(
    ptrdiff_t batch_size, uint64_t const *alphas, ptrdiff_t alphas_stride,
    ptrdiff_t *indices, ptrdiff_t indices_stride, void const *private_data);
typedef void (*ls_hs_internal_is_representative_kernel_type)(
    ptrdiff_t batch_size, uint64_t const *alphas, ptrdiff_t alphas_stride,
    uint8_t *are_representatives, double *norms, void const *private_data);
typedef void (*ls_hs_internal_state_info_kernel_type)(
    ptrdiff_t batch_size, uint64_t const *alphas, ptrdiff_t alphas_stride,
    uint64_t *betas, ptrdiff_t betas_stride, ls_hs_scalar *characters,
    double *norms, void const *private_data);

/* :100-107 */
typedef struct ls_hs_basis_kernels {
  ls_hs_internal_state_info_kernel_type state_info_kernel;
  void *state_info_data;
  ls_hs_internal_is_representative_kernel_type is_representative_kernel;
  void *is_representative_data;
  ls_hs_internal_state_index_kernel_type state_index_kernel;
  void *state_index_data;
} ls_hs_basis_kernels;

/* :109-119 -- refcount is `_Atomic int` in C11; same size/alignment as int. */
typedef struct ls_hs_permutation_group {
  int refcount;
  int number_bits;
  int number_shifts;
  int number_masks;
  uint64_t *masks;    /* [number_shifts][number_masks] row-major */
  uint64_t *shifts;   /* [number_shifts] */
  double *eigvals_re; /* [number_masks] */
  double *eigvals_im; /* [number_masks] */
  void *haskell_payload;
} ls_hs_permutation_group;

/* :121-133 */
typedef struct ls_hs_basis {
  int refcount;
  int number_sites;
  int number_particles; /* -1 when unset */
  int number_up;        /* -1 when unset */
  ls_hs_particle_type particle_type;
  int spin_inversion; /* 0 when none */
  bool state_index_is_identity;
  bool requires_projection;
  ls_hs_basis_kernels *kernels;
  chpl_external_array representatives;
  void *haskell_payload;
} ls_hs_basis;

/* :140-151 -- single 64-bit word per mask on this path */
typedef struct ls_hs_nonbranching_terms {
  int number_terms;
  int number_bits;
  ls_hs_scalar const *v;
  uint64_t const *m;
  uint64_t const *l;
  uint64_t const *r;
  uint64_t const *x;
  uint64_t const *s;
} ls_hs_nonbranching_terms;

/* :153-161 */
typedef struct ls_hs_operator {
  int refcount;
  ls_hs_basis const *basis;
  ls_hs_nonbranching_terms const *off_diag_terms; /* NULL when empty */
  ls_hs_nonbranching_terms const *diag_terms;     /* NULL when empty */
  void *haskell_payload;
} ls_hs_operator;

/* :170-180 -- the vtable the Chapel library registers at init
 * (chapel/src/LatticeSymmetries.chpl:18-33). */
typedef struct ls_chpl_kernels {
  void (*enumerate_states)(ls_hs_basis const *, uint64_t, uint64_t,
                           chpl_external_array *);
  void (*operator_apply_off_diag)(ls_hs_operator *, int64_t, uint64_t *,
                                  chpl_external_array *, chpl_external_array *,
                                  chpl_external_array *, int64_t);
  void (*operator_apply_diag)(ls_hs_operator *, int64_t, uint64_t *,
                              chpl_external_array *, int64_t);
  void (*matrix_vector_product)(ls_hs_operator *, int, double const *,
                                double *);
} ls_chpl_kernels;

/* ---- (1) replaces libkernels.a ------------------------------------------ */

/* kernels/reference.c:11-36 */
void ls_hs_set_exception_handler(void (*handler)(char const *message));
void ls_hs_error(char const *message);
void ls_hs_fatal_error(char const *func, int line, char const *message);

/* kernels/reference.c:40-64 */
void ls_hs_internal_destroy_external_array(chpl_external_array *arr);
int ls_hs_internal_read_refcount(int const *refcount);
void ls_hs_internal_write_refcount(int *refcount, int value);
int ls_hs_internal_inc_refcount(int *refcount);
int ls_hs_internal_dec_refcount(int *refcount);

/* kernels/kernels.c:112-193 -- copies the group tables to the device (the
 * reference borrows the Haskell-owned arrays); tolerates number_masks == 0. */
void *ls_internal_create_halide_kernel_data(ls_hs_permutation_group const *g,
                                            int spin_inversion);
void ls_internal_destroy_halide_kernel_data(void *p);

/* kernels/kernels.c:195-244, :246-319 -- names kept because the Haskell host
 * stores their ADDRESSES (haskell/src/LatticeSymmetries/Basis.hs:915-925). */
void ls_hs_is_representative_halide_kernel(ptrdiff_t batch_size,
                                           uint64_t const *alphas,
                                           ptrdiff_t alphas_stride,
                                           uint8_t *are_representatives,
                                           double *norms,
                                           void const *private_data);
void ls_hs_state_info_halide_kernel(ptrdiff_t batch_size,
                                    uint64_t const *alphas,
                                    ptrdiff_t alphas_stride, uint64_t *betas,
                                    ptrdiff_t betas_stride,
                                    ls_hs_scalar *characters, double *norms,
                                    void const *private_data);

/* kernels/indexing.c:94-127, :273-325 */
typedef struct ls_hs_state_index_binary_search_data
    ls_hs_state_index_binary_search_data;
ls_hs_state_index_binary_search_data *
ls_hs_create_state_index_binary_search_kernel_data(
    chpl_external_array const *representatives, int number_bits,
    int prefix_bits);
void ls_hs_destroy_state_index_binary_search_kernel_data(
    ls_hs_state_index_binary_search_data *cache);
void ls_hs_state_index_binary_search_kernel(ptrdiff_t batch_size,
                                            uint64_t const *spins,
                                            ptrdiff_t spins_stride,
                                            ptrdiff_t *indices,
                                            ptrdiff_t indices_stride,
                                            void const *private_kernel_data);

/* kernels/reference.c:137-211 */
void ls_hs_state_index(ls_hs_basis const *basis, ptrdiff_t batch_size,
                       uint64_t const *spins, ptrdiff_t spins_stride,
                       ptrdiff_t *indices, ptrdiff_t indices_stride);
void ls_hs_is_representative(ls_hs_basis const *basis, ptrdiff_t batch_size,
                             uint64_t const *alphas, ptrdiff_t alphas_stride,
                             uint8_t *are_representatives, double *norms);
void ls_hs_state_info(ls_hs_basis const *basis, ptrdiff_t batch_size,
                      uint64_t const *alphas, ptrdiff_t alphas_stride,
                      uint64_t *betas, ptrdiff_t betas_stride,
                      ls_hs_scalar *characters, double *norms);
void ls_hs_build_representatives(ls_hs_basis *basis, uint64_t lower,
                                 uint64_t upper);
void ls_hs_unchecked_set_representatives(ls_hs_basis *basis,
                                         chpl_external_array const *states,
                                         int cache_bits);

/* kernels/reference.c:67-134 */
void ls_internal_operator_apply_diag_x1(ls_hs_operator const *op,
                                        ptrdiff_t batch_size,
                                        uint64_t const *alphas, double *ys,
                                        double const *xs);
void ls_internal_operator_apply_off_diag_x1(ls_hs_operator const *op,
                                            ptrdiff_t batch_size,
                                            uint64_t const *alphas,
                                            uint64_t *betas,
                                            ls_hs_scalar *coeffs,
                                            ptrdiff_t *offsets,
                                            double const *xs);

/* kernels/reference.c:214-225 */
ls_chpl_kernels const *ls_hs_internal_get_chpl_kernels(void);
void ls_hs_internal_set_chpl_kernels(ls_chpl_kernels const *kernels);

/* ---- (2) replaces liblattice_symmetries_chapel.so ------------------------ */

/* chapel/src/library.c:19-34; LatticeSymmetries.chpl:29-33 */
void ls_chpl_init(void);
void ls_chpl_finalize(void);
void ls_chpl_init_kernels(void);

/* chapel/src/StatesEnumeration.chpl:692-709 (lower/upper are ignored there
 * and recomputed from the basis, :683-688; same here). */
void ls_chpl_enumerate_representatives(ls_hs_basis const *basis,
                                       uint64_t lower, uint64_t upper,
                                       chpl_external_array *dest);
/* chapel/src/BatchedOperator.chpl:298-357 */
void ls_chpl_operator_apply_diag(ls_hs_operator *op, int64_t count,
                                 uint64_t *alphas, chpl_external_array *coeffs,
                                 int64_t num_tasks);
void ls_chpl_operator_apply_off_diag(ls_hs_operator *op, int64_t count,
                                     uint64_t *alphas,
                                     chpl_external_array *betas,
                                     chpl_external_array *coeffs,
                                     chpl_external_array *offsets,
                                     int64_t num_tasks);
/* chapel/src/DistributedMatrixVector.chpl:1090-1105 */
void ls_chpl_matrix_vector_product(ls_hs_operator *op, int num_vectors,
                                   double const *x, double *y);

/* ---- extensions: device-resident path ----------------------------------- */

/* Number of CUDA kernels this library has launched in this process. */
uint64_t ls_b200_kernel_launch_count(void);
/* Device-time (ms, CUDA events on the library stream) of the most recent
 * matvec / basis-build kernel; name selects "matvec" or "build". */
double ls_b200_last_kernel_ms(char const *name);
/* The CUDA stream (cudaStream_t) all library work is ordered on. */
void *ls_b200_stream(void);
/* Order all subsequent library work on the caller's stream instead (e.g. the
 * stream NCCL collectives are enqueued behind); NULL restores the library's
 * own stream.  Drains the previous stream first.  Returns 0 on success. */
int ls_b200_set_stream(void *cuda_stream);
int ls_b200_device_count(void);
/* Measured 32-bit logic-op (LOP3) issue rate of the device, thread-ops per
 * second: the integer-side roofline denominator (MEASURED_PEAKS.json has only
 * HBM and tensor peaks). */
double ls_b200_measure_lop3_peak(void);

/* Device views of a built basis: sorted representatives, their norms
 * (state_info convention, sqrt(n/|G|)); pointers stay owned by the library. */
int ls_b200_basis_device_view(ls_hs_basis const *basis,
                              uint64_t const **representatives,
                              double const **norms, uint64_t *count);

/* Shape of the state -> index structure that replaces
 * ls_hs_state_index_binary_search_data (kernels/indexing.c:10-18):
 * out[0] = prefix bits of the first-level table, out[1] = trip count of the
 * final branchless search (bit length of the widest window), out[2] = 1 when
 * crowded buckets carry second-level tables, out[3] = number of states.
 * Returns 0, or -1 when the basis is not built. */
int ls_b200_index_info(ls_hs_basis const *basis, int64_t out[4]);

/* y[row_begin:row_end] = (H x)[row_begin:row_end] with x (full length dim)
 * and y (length row_end-row_begin) in DEVICE memory; asynchronous on
 * ls_b200_stream().  Returns 0 on success.  The rows are the contiguous
 * representative range owned by this rank (multi-GPU: x is the all-gathered
 * vector). */
int ls_b200_matvec_device(ls_hs_operator const *op, int64_t row_begin,
                          int64_t row_end, double const *x_dev, double *y_dev);
/* Complex128 variant (interleaved re, im).  Extension: the reference's matvec
 * is real-only (DistributedMatrixVector.chpl:1090-1091). */
/* The product in two phases (split path; other paths do everything in phase 2):
 * phase 1 canonicalises every matrix element of the row range -- none of that
 * depends on x, so x_dev / y_dev may be NULL -- into buffers that persist;
 * phase 2 ranks, gathers and sums.  A multi-GPU caller overlaps phase 1 of the
 * next product with the NCCL all-gather of this product's result.  Every
 * product still performs both phases exactly once. */
int ls_b200_matvec_device_phase(ls_hs_operator const *op, int64_t row_begin,
                                int64_t row_end, void const *x_dev, void *y_dev,
                                int complex_vectors, int phase);
/* Block matvec (extension; the reference halts on numVectors != 1,
 * chapel/src/DistributedMatrixVector.chpl:1096-1097): vector v is
 * x_dev + v x_stride -> y_dev + v y_stride (strides in doubles).  All vectors
 * share ONE canonicalisation + ranking pass over the matrix elements, so k
 * vectors cost far less than k products.  ls_chpl_matrix_vector_product accepts
 * num_vectors > 1 the same way (contiguous vectors of length dim). */
int ls_b200_matvec_block_device(ls_hs_operator const *op, int64_t row_begin,
                                int64_t row_end, int number_vectors,
                                double const *x_dev, int64_t x_stride,
                                double *y_dev, int64_t y_stride);
int ls_b200_matvec_device_c128(ls_hs_operator const *op, int64_t row_begin,
                               int64_t row_end, ls_hs_scalar const *x_dev,
                               ls_hs_scalar *y_dev);
/* Waits for the device matvecs issued so far; returns 0 on success, 1 (after
 * calling ls_hs_error) when an operator left the basis with a non-zero
 * coefficient -- the reference's "invalid index" halt,
 * chapel/src/DistributedMatrixVector.chpl:127-135. */
int ls_b200_matvec_sync(void);
/* Number of off-diagonal matrix elements (emitted (row, term) pairs) in the
 * row range: the unit of the matvec throughput metric. */
int64_t ls_b200_count_matrix_elements(ls_hs_operator const *op,
                                      int64_t row_begin, int64_t row_end);

/* Sharded basis construction: scan the candidates whose combinadic / linear
 * index lies in [index_begin, index_end) of the basis' enumeration range and
 * return this shard's representatives (ascending) and norms in device memory
 * (caller frees with ls_b200_device_free).  total_candidates (optional out)
 * is the size of the whole enumeration range. */
int ls_b200_build_shard(ls_hs_basis const *basis, uint64_t index_begin,
                        uint64_t index_end, uint64_t **representatives_dev,
                        double **norms_dev, uint64_t *count);
/* The same for a rank's whole block-cyclic share in one call: blocks
 * [first_begin + k stride, first_begin + k stride + block_size), k = 0 ..
 * number_blocks - 1 (clipped to the enumeration range); the output is their
 * concatenation, block_counts[k] (caller-provided) the states found in block k. */
int ls_b200_build_blocks(ls_hs_basis const *basis, uint64_t first_begin,
                         uint64_t block_size, uint64_t stride,
                         uint64_t number_blocks, uint64_t **representatives_dev,
                         double **norms_dev, uint64_t *block_counts);
uint64_t ls_b200_number_candidates(ls_hs_basis const *basis);
/* Install an (all-gathered) device-resident representative list + norms as
 * the basis' representatives; the library takes ownership of both buffers.
 * basis->representatives.elts (a host pointer in the reference) is served by a
 * managed allocation that doubles as the device array (read-mostly: host reads
 * fault in duplicates, nothing is copied up front); LS_B200_HOST_MIRROR=pinned
 * keeps a separate pinned host copy instead. */
int ls_b200_set_representatives_device(ls_hs_basis *basis,
                                       uint64_t *representatives_dev,
                                       double *norms_dev, uint64_t count,
                                       int cache_bits);
void *ls_b200_device_malloc(size_t bytes);
void ls_b200_device_free(void *p);
/* Copies ordered on ls_b200_stream(); both return after the copy completed
 * (0 on success).  `host` may be pageable or pinned. */
int ls_b200_copy_to_device(void *dst_dev, void const *src_host, size_t bytes);
int ls_b200_copy_to_host(void *dst_host, void const *src_dev, size_t bytes);
/* Pinned host allocations for callers that stage vectors themselves. */
void *ls_b200_host_malloc(size_t bytes);
void ls_b200_host_free(void *p);

/* Extension (rows of the symmetry-PROJECTED operator, batched; semantics of chapel/src/BatchedOperator.chpl:207-253):
 * for every alphas[i], the entries offsets[i] .. offsets[i+1] hold the representative of each image, and the matrix
 * element chi c n(beta) / n(alpha_i) (0 for images of vanishing norm).  Host pointers; reps / coeffs / indices need
 * room for count * (number of off-diagonal terms) entries, offsets for count + 1.  indices (may be NULL): position of
 * each representative in the built basis, -1 when absent.  Returns the number of entries, -1 on error. */
int64_t ls_b200_operator_apply_off_diag_projected(ls_hs_operator const *op, int64_t count, uint64_t const *alphas,
                                                  uint64_t *reps, ls_hs_scalar *coeffs, int64_t *offsets,
                                                  int64_t *indices);

/* Drops the device-side tables cached for an operator (term tables, phased-product buffers); call before the
 * ls_hs_operator struct is freed.  Harmless for operators the library has never seen. */
void ls_b200_operator_release(ls_hs_operator const *op);

/* ---- extensions: several GPUs of one node ------------------------------------------
 * Replaces the multi-locale half of the Chapel library (chapel/src/StatesEnumeration.chpl:198-212, :537-602;
 * chapel/src/DistributedMatrixVector.chpl:179-339, :545-579, :775-807, :1060-1088).  One process per GPU; the
 * library owns its NCCL communicator: rank 0 calls ls_b200_comm_unique_id, the host distributes the 128 bytes by
 * any means (MPI, torch.distributed, a file), every rank calls ls_b200_comm_init.  From then on
 *   ls_hs_build_representatives      builds the basis across all ranks; basis->representatives is THIS rank's
 *                                    block: a contiguous range of the globally sorted list (not a hash class);
 *   ls_chpl_matrix_vector_product    takes this rank's blocks of x and y (host pointers), like the per-locale
 *                                    blocks of DistributedMatrixVector.chpl:1066-1073.
 * Every rank must make the same calls in the same order (they contain collectives). */
int ls_b200_comm_unique_id(void *out, size_t bytes); /* bytes >= 128 */
int ls_b200_comm_init(int world, int rank, void const *unique_id);
void ls_b200_comm_finalize(void);
int ls_b200_comm_size(void);
int ls_b200_comm_rank(void);
/* In-place sum over the ranks of `count` doubles in device memory, ordered on ls_b200_stream() (Lanczos dots). */
int ls_b200_comm_allreduce_f64(double *values_dev, int count);
/* The sharded build, explicitly.  balance_for (optional): move the row boundaries so that every rank holds the same
 * number of matrix elements of that operator.  flags: 1 = no replicated index (all-to-all products only),
 * 2 = replicate compact keys only ("wide" index) even for small bases, 4 = keep the even row split. */
int ls_b200_dist_build(ls_hs_basis *basis, ls_hs_operator const *balance_for, int flags);
/* y_local = (H x)_local with this rank's blocks of x and y in DEVICE memory; asynchronous on ls_b200_stream()
 * (finish with ls_b200_matvec_sync).  mode 0 = automatic, 1 = all-gather form (x replicated: compact keys of the
 * whole basis on every rank, pre-scaled vector all-gathered in place, pull-form kernels), 2 = all-to-all form
 * ((representative, coefficient) records grouped by owner rank, one grouped send/recv exchange per chunk of
 * columns, the owner ranks locally and adds with fp64 atomics). */
int ls_b200_dist_matvec(ls_hs_operator const *op, double const *x_local_dev, double *y_local_dev, int mode);
int ls_b200_dist_matvec_c128(ls_hs_operator const *op, ls_hs_scalar const *x_local_dev, ls_hs_scalar *y_local_dev,
                             int mode); /* all-gather form only */
/* Moves the row boundaries so that every rank spends the same TIME on a product, using the kernel times measured
 * during the LAST product on this basis (requires LS_B200_PROFILE=1 in the environment; rows at the high end of the
 * sorted list cost up to 2.6x more per matrix element than those at the low end because their lookups scatter).
 * Collective.  Afterwards the local block -- basis->representatives, ls_b200_dist_info -- has changed and the
 * caller's vectors must be re-split.  Returns 0 when rows moved, 1 when nothing changed, -1 on error. */
int ls_b200_dist_rebalance(ls_hs_basis *basis);
/* The same for virtual ranks with explicit costs: costs[r * 1024 + k] = cost of the k-th of 1024 equal pieces of
 * virtual rank r's rows. */
int ls_b200_emu_rebalance(ls_hs_basis **bases, int world, double const *costs);
/* out = {world, rank, dim over all ranks, first row, one past the last row of this rank,
 *        replicated index (0 none, 1 two-level, 2 wide), its search trip count, its prefix bits}; -1 when the basis
 * is not sharded. */
int ls_b200_dist_info(ls_hs_basis const *basis, int64_t out[8]);
/* Row boundaries of all ranks (world + 1 values); returns world, or -1. */
int ls_b200_dist_bounds(ls_hs_basis const *basis, int64_t *bounds, int capacity);
/* The same drivers for `world` VIRTUAL ranks on the current device, run in lockstep with device-to-device copies in
 * place of the NCCL collectives (bases[r] / ops[r] / x_dev[r] / y_dev[r] belong to virtual rank r).  Exercises the
 * sharded build, both product forms and the wide index on a single GPU. */
int ls_b200_emu_build(ls_hs_basis **bases, ls_hs_operator const *const *balance_for, int world, int flags);
int ls_b200_emu_matvec(ls_hs_operator const *const *ops, int world, double const *const *x_dev, double *const *y_dev,
                       int mode, int complex_vectors);
/* Host-side planning of the sharded build (pure functions, no device): the block plan of the candidate range, the
 * all-to-all-v plan that turns block-cyclic pieces into contiguous row ranges, cost-balanced row boundaries. */
int64_t ls_b200_plan_blocks(uint64_t total, int world, uint64_t *begins, uint64_t *ends, int64_t capacity);
int64_t ls_b200_plan_redistribution(int world, int me, int64_t number_pieces, int64_t const *lengths,
                                    int32_t const *owners, int64_t const *bounds, int64_t *scount, int64_t *sdispl,
                                    int64_t *rcount, int64_t *rdispl, int64_t *places, int64_t capacity);
int ls_b200_plan_balanced_bounds(int64_t number_blocks, int64_t const *edges, double const *costs, int world,
                                 int64_t *bounds);
/* ... of equal cost with at most `cap` rows per rank, boundaries interpolated inside a block (ls_b200_dist_rebalance). */
int ls_b200_plan_rebalance_bounds(int64_t number_blocks, int64_t const *edges, double const *costs, int world, int64_t cap,
                                  int64_t *bounds);
/* A rank's block-cyclic share of the candidate range: out = {block shift (a block holds 32 << shift candidates), blocks
 * over all ranks, blocks of this rank, candidates of this rank}. */
int ls_b200_plan_cyclic_share(uint64_t total, int world, int rank, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
