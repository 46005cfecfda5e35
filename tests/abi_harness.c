/* abi_harness.c -- a C caller that knows ONLY the reference's own header (kernels/lattice_symmetries_types.h, compiled
 * with -I/root/reference/kernels -DLS_NO_STD_COMPLEX) and links liblattice_symmetries_b200.so: the drop-in boundary seen
 * from the reference's side.  Fills the structs the way haskell/src/LatticeSymmetries/FFI.hs:92-143 lays them out.
 *
 *   abi_harness layout                  print sizeof / offsetof of every struct of the REFERENCE header (no device)
 *   abi_harness run problem.bin out.bin build the basis (ls_hs_build_representatives), one y = H x through
 *                                       ls_hs_internal_get_chpl_kernels()->matrix_vector_product; representatives and y
 *                                       go to out.bin
 * problem.bin (written by tests/test_abi.py): int32 number_sites, hamming_weight, spin_inversion, T_off, T_diag, then
 * for each table v[T] (re, im doubles), m, l, r, x, s (uint64 each), then dim-agnostic seed for x.                 */
#include "lattice_symmetries_types.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* entry points the Haskell host / Chapel / Python import (declared in the reference's sources, not in this header) */
void ls_hs_build_representatives(ls_hs_basis *basis, uint64_t lower, uint64_t upper);
void ls_chpl_init(void);

#define FIELD(T, f) printf("\"%s.%s\": %zu, ", #T, #f, offsetof(T, f))
#define SIZE(T) printf("\"sizeof %s\": %zu, ", #T, sizeof(T))

static void print_layout(void) {
  printf("{");
  SIZE(chpl_external_array); SIZE(ls_hs_scalar); SIZE(ls_hs_basis_kernels); SIZE(ls_hs_permutation_group);
  SIZE(ls_hs_basis); SIZE(ls_hs_nonbranching_terms); SIZE(ls_hs_operator); SIZE(ls_chpl_kernels);
  FIELD(chpl_external_array, elts); FIELD(chpl_external_array, num_elts); FIELD(chpl_external_array, freer);
  FIELD(ls_hs_basis_kernels, state_info_kernel); FIELD(ls_hs_basis_kernels, state_info_data);
  FIELD(ls_hs_basis_kernels, is_representative_kernel); FIELD(ls_hs_basis_kernels, is_representative_data);
  FIELD(ls_hs_basis_kernels, state_index_kernel); FIELD(ls_hs_basis_kernels, state_index_data);
  FIELD(ls_hs_permutation_group, refcount); FIELD(ls_hs_permutation_group, number_bits);
  FIELD(ls_hs_permutation_group, number_shifts); FIELD(ls_hs_permutation_group, number_masks);
  FIELD(ls_hs_permutation_group, masks); FIELD(ls_hs_permutation_group, shifts);
  FIELD(ls_hs_permutation_group, eigvals_re); FIELD(ls_hs_permutation_group, eigvals_im);
  FIELD(ls_hs_permutation_group, haskell_payload);
  FIELD(ls_hs_basis, refcount); FIELD(ls_hs_basis, number_sites); FIELD(ls_hs_basis, number_particles);
  FIELD(ls_hs_basis, number_up); FIELD(ls_hs_basis, particle_type); FIELD(ls_hs_basis, spin_inversion);
  FIELD(ls_hs_basis, state_index_is_identity); FIELD(ls_hs_basis, requires_projection); FIELD(ls_hs_basis, kernels);
  FIELD(ls_hs_basis, representatives); FIELD(ls_hs_basis, haskell_payload);
  FIELD(ls_hs_nonbranching_terms, number_terms); FIELD(ls_hs_nonbranching_terms, number_bits);
  FIELD(ls_hs_nonbranching_terms, v); FIELD(ls_hs_nonbranching_terms, m); FIELD(ls_hs_nonbranching_terms, l);
  FIELD(ls_hs_nonbranching_terms, r); FIELD(ls_hs_nonbranching_terms, x); FIELD(ls_hs_nonbranching_terms, s);
  FIELD(ls_hs_operator, refcount); FIELD(ls_hs_operator, basis); FIELD(ls_hs_operator, off_diag_terms);
  FIELD(ls_hs_operator, diag_terms); FIELD(ls_hs_operator, haskell_payload);
  FIELD(ls_chpl_kernels, enumerate_states); FIELD(ls_chpl_kernels, operator_apply_off_diag);
  FIELD(ls_chpl_kernels, operator_apply_diag); FIELD(ls_chpl_kernels, matrix_vector_product);
  printf("\"LS_HS_SPIN\": %d, \"LS_HS_SPINFUL_FERMION\": %d, \"LS_HS_SPINLESS_FERMION\": %d}\n", (int)LS_HS_SPIN,
         (int)LS_HS_SPINFUL_FERMION, (int)LS_HS_SPINLESS_FERMION);
}

static ls_hs_nonbranching_terms *read_terms(FILE *f, int T, int bits) {
  if (T == 0) return NULL;
  ls_hs_nonbranching_terms *t = calloc(1, sizeof *t);
  uint64_t *cols = malloc(sizeof(uint64_t) * 5 * (size_t)T);
  ls_hs_scalar *v = malloc(sizeof(ls_hs_scalar) * (size_t)T);
  if (fread(v, sizeof(ls_hs_scalar), (size_t)T, f) != (size_t)T || fread(cols, 8, 5 * (size_t)T, f) != 5 * (size_t)T) exit(3);
  t->number_terms = T; t->number_bits = bits; t->v = v;
  t->m = cols; t->l = cols + T; t->r = cols + 2 * T; t->x = cols + 3 * T; t->s = cols + 4 * T;
  return t;
}

int main(int argc, char **argv) {
  if (argc >= 2 && strcmp(argv[1], "layout") == 0) { print_layout(); return 0; }
  if (argc < 4 || strcmp(argv[1], "run") != 0) { fprintf(stderr, "usage: abi_harness layout | run in out\n"); return 2; }
  FILE *f = fopen(argv[2], "rb");
  int32_t head[5];
  if (!f || fread(head, 4, 5, f) != 5) return 3;
  int const n = head[0], hw = head[1], inv = head[2];
  ls_hs_basis_kernels kernels; memset(&kernels, 0, sizeof kernels);      /* no permutation group: all kernels NULL */
  ls_hs_basis basis; memset(&basis, 0, sizeof basis);
  basis.refcount = 1; basis.number_sites = n; basis.number_particles = n; basis.number_up = hw;
  basis.particle_type = LS_HS_SPIN; basis.spin_inversion = inv; basis.state_index_is_identity = false;
  basis.requires_projection = inv != 0; basis.kernels = &kernels;
  ls_hs_operator op; memset(&op, 0, sizeof op);
  op.refcount = 1; op.basis = &basis;
  op.off_diag_terms = read_terms(f, head[3], n); op.diag_terms = read_terms(f, head[4], n);
  fclose(f);
  ls_chpl_init();                                                       /* registers the four vtable entries */
  uint64_t const lo = hw > 0 ? (((uint64_t)1 << hw) - 1) : 0;           /* Basis.hs:701-740 min / max state estimates */
  uint64_t hi = lo << (n - hw);
  if (inv != 0) hi = hw == 0 ? 0 : lo << (n - hw - 1);                  /* top bit clear: the smaller of (s, ~s) */
  ls_hs_build_representatives(&basis, lo, hi);
  uint64_t const dim = basis.representatives.num_elts;
  double *x = malloc(8 * dim), *y = malloc(8 * dim);
  for (uint64_t i = 0; i < dim; ++i) x[i] = 0.5 + (double)((i * 2654435761u) % 1000) / 1000.0;
  ls_hs_internal_get_chpl_kernels()->matrix_vector_product(&op, 1, x, y);
  FILE *o = fopen(argv[3], "wb");
  fwrite(&dim, 8, 1, o); fwrite(basis.representatives.elts, 8, dim, o); fwrite(x, 8, dim, o); fwrite(y, 8, dim, o);
  fclose(o);
  ls_hs_internal_destroy_external_array(&basis.representatives);
  printf("abi_harness: dim %llu\n", (unsigned long long)dim);
  return 0;
}
