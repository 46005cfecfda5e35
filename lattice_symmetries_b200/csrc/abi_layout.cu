// abi_layout.cu -- compile-time pins of the struct layouts the Haskell host
// marshals into (haskell/src/LatticeSymmetries/FFI.hs:69-143, generic-derived
// Storable in declaration order) and the Python cffi wrapper reads
// (kernels/lattice_symmetries_types.h, x86-64 SysV / LP64).
#include <cstddef>

#include "../../include/lattice_symmetries_b200.h"

#define PIN(type, field, off) static_assert(offsetof(type, field) == (off), #type "." #field " offset")

// kernels/lattice_symmetries_types.h:27-32
PIN(chpl_external_array, elts, 0);
PIN(chpl_external_array, num_elts, 8);
PIN(chpl_external_array, freer, 16);
static_assert(sizeof(chpl_external_array) == 24, "chpl_external_array size");

// :43-56
static_assert(sizeof(ls_hs_scalar) == 16 && alignof(ls_hs_scalar) == 8, "ls_hs_scalar is a C99 double complex");

// :100-107
PIN(ls_hs_basis_kernels, state_info_kernel, 0);
PIN(ls_hs_basis_kernels, state_info_data, 8);
PIN(ls_hs_basis_kernels, is_representative_kernel, 16);
PIN(ls_hs_basis_kernels, is_representative_data, 24);
PIN(ls_hs_basis_kernels, state_index_kernel, 32);
PIN(ls_hs_basis_kernels, state_index_data, 40);
static_assert(sizeof(ls_hs_basis_kernels) == 48, "ls_hs_basis_kernels size");

// :109-119
PIN(ls_hs_permutation_group, refcount, 0);
PIN(ls_hs_permutation_group, number_bits, 4);
PIN(ls_hs_permutation_group, number_shifts, 8);
PIN(ls_hs_permutation_group, number_masks, 12);
PIN(ls_hs_permutation_group, masks, 16);
PIN(ls_hs_permutation_group, shifts, 24);
PIN(ls_hs_permutation_group, eigvals_re, 32);
PIN(ls_hs_permutation_group, eigvals_im, 40);
PIN(ls_hs_permutation_group, haskell_payload, 48);
static_assert(sizeof(ls_hs_permutation_group) == 56, "ls_hs_permutation_group size");

// :121-133
PIN(ls_hs_basis, refcount, 0);
PIN(ls_hs_basis, number_sites, 4);
PIN(ls_hs_basis, number_particles, 8);
PIN(ls_hs_basis, number_up, 12);
PIN(ls_hs_basis, particle_type, 16);
PIN(ls_hs_basis, spin_inversion, 20);
PIN(ls_hs_basis, state_index_is_identity, 24);
PIN(ls_hs_basis, requires_projection, 25);
PIN(ls_hs_basis, kernels, 32);
PIN(ls_hs_basis, representatives, 40);
PIN(ls_hs_basis, haskell_payload, 64);
static_assert(sizeof(ls_hs_basis) == 72, "ls_hs_basis size");
static_assert(sizeof(ls_hs_particle_type) == 4, "enum is int-sized");

// :140-151
PIN(ls_hs_nonbranching_terms, number_terms, 0);
PIN(ls_hs_nonbranching_terms, number_bits, 4);
PIN(ls_hs_nonbranching_terms, v, 8);
PIN(ls_hs_nonbranching_terms, m, 16);
PIN(ls_hs_nonbranching_terms, l, 24);
PIN(ls_hs_nonbranching_terms, r, 32);
PIN(ls_hs_nonbranching_terms, x, 40);
PIN(ls_hs_nonbranching_terms, s, 48);
static_assert(sizeof(ls_hs_nonbranching_terms) == 56, "ls_hs_nonbranching_terms size");

// :153-161
PIN(ls_hs_operator, refcount, 0);
PIN(ls_hs_operator, basis, 8);
PIN(ls_hs_operator, off_diag_terms, 16);
PIN(ls_hs_operator, diag_terms, 24);
PIN(ls_hs_operator, haskell_payload, 32);
static_assert(sizeof(ls_hs_operator) == 40, "ls_hs_operator size");

// :170-180
PIN(ls_chpl_kernels, enumerate_states, 0);
PIN(ls_chpl_kernels, operator_apply_off_diag, 8);
PIN(ls_chpl_kernels, operator_apply_diag, 16);
PIN(ls_chpl_kernels, matrix_vector_product, 24);
static_assert(sizeof(ls_chpl_kernels) == 32, "ls_chpl_kernels size");
