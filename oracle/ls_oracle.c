/*
 * oracle/ls_oracle.c -- CPU restatement of the lattice-symmetries hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lattice_symmetries_b200/ (the
 * product) may import, link or execute this file.  It is used by tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py, as the checker and as the timed CPU baseline -- never as the thing
 * shipped.
 *
 * Every function cites the reference file:line (tree: twesterhout/
 * lattice-symmetries @ 36215fe) whose arithmetic it restates.  The Halide
 * generators (kernels/generator.cpp) cannot be compiled here (Halide 14.0.0 is
 * an un-vendored third-party toolchain), so state_info / is_representative are
 * restated from the generator source; kernels/indexing.c and
 * kernels/reference.c DO compile and are built into oracle/_ref/libref.so by
 * oracle/Makefile -- tests/test_oracle.py pins this restatement against them.
 *
 * Parity pinning: see tests/test_oracle.py (reference known answers: Benes
 * vectors, basis lists, Hubbard dense matrices, E0 = -18.06178542, the five
 * HPhi energies) and oracle/_ref (compiled reference C).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LS_ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* kernels/generator.cpp:5-8  bit_permute_step_64                             */
static inline uint64_t bit_permute_step(uint64_t x, uint64_t m, uint64_t d) {
  uint64_t const y = ((x >> d) ^ x) & m;
  return (x ^ y) ^ (y << d);
}

/* kernels/kernels.c:96-98  get_flip_mask_64 */
static inline uint64_t flip_mask_64(unsigned n) {
  return n == 0U ? (uint64_t)0 : ((~(uint64_t)0) >> (64U - n));
}

/* Group tables, laid out exactly as ls_hs_permutation_group
 * (kernels/lattice_symmetries_types.h:109-119; Benes.hs:356-377):
 *   masks  u64[depth][number_masks] row-major, shifts u64[depth],
 *   eigvals_re/im f64[number_masks]. */
typedef struct oracle_group {
  int number_bits;
  int depth;
  int number_masks;
  int spin_inversion; /* 0, +1, -1 */
  uint64_t const *masks;
  uint64_t const *shifts;
  double const *eigvals_re;
  double const *eigvals_im;
} oracle_group;

#define ORACLE_MAX_GROUP 4096

/* Apply every group element's Benes network to x (generator.cpp:88-94). The
 * loop nest is [stage][element] so that gcc vectorises over elements, the same
 * axis Halide vectorises (generator.cpp:157-158). */
static inline void orbit_images(oracle_group const *g, uint64_t x,
                                uint64_t *restrict y) {
  int const G = g->number_masks;
  for (int j = 0; j < G; ++j) y[j] = x;
  for (int k = 0; k < g->depth; ++k) {
    uint64_t const d = g->shifts[k];
    uint64_t const *restrict m = g->masks + (size_t)k * (size_t)G;
    for (int j = 0; j < G; ++j) y[j] = bit_permute_step(y[j], m[j], d);
  }
}

/* kernels/generator.cpp:24-54 (reduction_step_impl / reduction_step) and
 * :77-140 (generate): running tuple (r, c_re, c_im, n) initialised to
 * (x, 1, 0, 0); y < r replaces (r, c); y == x adds Re chi to n; with spin
 * inversion the step repeats with y ^ flip_mask and +-chi.  The sum is taken
 * sequentially in group order (identity first, Group.hs:179) -- Halide's
 * lane-wise order (generator.cpp:103-127) is CPU-SIMD-width dependent and only
 * differs for mathematically-zero norms in complex sectors. */
LS_ORACLE_API void oracle_state_info(oracle_group const *g, ptrdiff_t batch,
                                     uint64_t const *alphas, uint64_t *betas,
                                     double *characters /* [batch][2] */,
                                     double *norms) {
  int const G = g->number_masks;
  uint64_t const flip = flip_mask_64((unsigned)g->number_bits);
  int const inv = g->spin_inversion;
#pragma omp parallel
  {
    uint64_t *y = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(G > 0 ? G : 1));
#pragma omp for schedule(static)
    for (ptrdiff_t i = 0; i < batch; ++i) {
      uint64_t const x = alphas[i];
      orbit_images(g, x, y);
      uint64_t r = x;
      double c_re = 1.0, c_im = 0.0, n = 0.0;
      for (int j = 0; j < G; ++j) {
        double const re = g->eigvals_re[j], im = g->eigvals_im[j];
        if (y[j] < r) { r = y[j]; c_re = re; c_im = im; }
        if (y[j] == x) n += re;
        if (inv != 0) {
          uint64_t const yf = y[j] ^ flip;
          double const s = (double)inv;
          if (yf < r) { r = yf; c_re = s * re; c_im = s * im; }
          if (yf == x) n += s * re;
        }
      }
      betas[i] = r;
      characters[2 * i + 0] = c_re;
      characters[2 * i + 1] = c_im;
      /* generator.cpp:135-140 */
      norms[i] = sqrt(n / (double)((inv == 0 ? 1 : 2) * G));
    }
    free(y);
  }
}

/* kernels/generator.cpp:181-253: flag = AND_j (y_j >= x [&& y_j^flip >= x]);
 * s = sum_j ([y_j==x] +- [y_j^flip==x]) * Re chi_j; outputs norm = s (raw) and
 * is_representative = (s > 0) ? flag : 0.  No early stop here (the reference
 * stops per SIMD chunk, :236, leaving s partial when flag is 0), so compare
 * norms only where the flag is set. */
LS_ORACLE_API void oracle_is_representative(oracle_group const *g,
                                            ptrdiff_t batch,
                                            uint64_t const *alphas,
                                            uint8_t *flags, double *norms) {
  int const G = g->number_masks;
  uint64_t const flip = flip_mask_64((unsigned)g->number_bits);
  int const inv = g->spin_inversion;
#pragma omp parallel
  {
    uint64_t *y = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(G > 0 ? G : 1));
#pragma omp for schedule(static)
    for (ptrdiff_t i = 0; i < batch; ++i) {
      uint64_t const x = alphas[i];
      orbit_images(g, x, y);
      int flag = 1;
      double s = 0.0;
      for (int j = 0; j < G; ++j) {
        double t = (double)(y[j] == x);
        int ge = y[j] >= x;
        if (inv != 0) {
          uint64_t const yf = y[j] ^ flip;
          t = (inv == 1) ? t + (double)(yf == x) : t - (double)(yf == x);
          ge = ge && (yf >= x);
        }
        s += t * g->eigvals_re[j];
        flag = flag && ge;
      }
      norms[i] = s;
      flags[i] = (s > 0.0) ? (uint8_t)flag : (uint8_t)0;
    }
    free(y);
  }
}

/* ------------------------------------------------------------------------- */
/* Combinadics: haskell/src/LatticeSymmetries/Basis.hs:487-550               */
static uint64_t g_binom[65][65];
static int g_binom_ready = 0;
static void init_binomials(void) {
  if (g_binom_ready) return;
  memset(g_binom, 0, sizeof(g_binom));
  for (int n = 0; n <= 64; ++n) {
    g_binom[n][0] = 1;
    for (int k = 1; k <= n; ++k)
      g_binom[n][k] = g_binom[n - 1][k - 1] + (k <= n - 1 ? g_binom[n - 1][k] : 0);
  }
  g_binom_ready = 1;
}
/* Basis.hs:507-510: binomial n k = 0 when n <= 0 || k > n */
static inline uint64_t binomial(int n, int k) {
  if (n <= 0 || k > n) return 0;
  return g_binom[n][k];
}

/* Basis.hs:519-528 fixedHammingStateToIndex */
LS_ORACLE_API int64_t oracle_fixed_hamming_state_to_index(uint64_t alpha) {
  init_binomials();
  int64_t i = 0;
  int k = 1;
  while (alpha != 0) {
    int const c = __builtin_ctzll(alpha);
    alpha &= alpha - 1;
    i += (int64_t)binomial(c, k);
    ++k;
  }
  return i;
}

/* Basis.hs:530-550 fixedHammingIndexToState: for i = hw..1 pick the largest c
 * with binomial(c, i) <= index. */
LS_ORACLE_API uint64_t oracle_fixed_hamming_index_to_state(int64_t index,
                                                           int hamming_weight) {
  init_binomials();
  uint64_t state = 0;
  for (int i = hamming_weight; i > 0; --i) {
    int c = i - 1;
    uint64_t contribution = 0;
    while (c < 64) {
      uint64_t const next = binomial(c + 1, i);
      if (next > (uint64_t)index) break;
      ++c;
      contribution = next;
    }
    state |= (uint64_t)1 << c;
    index -= (int64_t)contribution;
  }
  return state;
}

/* chapel/src/StatesEnumeration.chpl:29-32 nextStateFixedHamming */
static inline uint64_t next_state_fixed_hamming(uint64_t v) {
  uint64_t const t = v | (v - 1);
  return (t + 1) | (((~t & (t + 1)) - 1) >> (__builtin_ctzll(v) + 1));
}

static int oracle_num_threads_impl(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

typedef struct oracle_vec {
  uint64_t *data;
  size_t size, cap;
} oracle_vec;
static void vec_push(oracle_vec *v, uint64_t x) {
  if (v->size == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 1024;
    v->data = (uint64_t *)realloc(v->data, v->cap * sizeof(uint64_t));
  }
  v->data[v->size++] = x;
}

/* Basis description mirroring ls_hs_basis (lattice_symmetries_types.h:121-133)
 * plus the predicates the Chapel side queries from Haskell
 * (Basis.hs:701-774). */
typedef struct oracle_basis {
  int number_sites;
  int number_particles; /* -1 when unset */
  int number_up;        /* -1 when unset */
  int particle_type;    /* 0 spin, 1 spinful fermion, 2 spinless fermion */
  int spin_inversion;   /* 0 when none */
  int has_permutation_symmetries;
  oracle_group group;   /* valid when has_permutation_symmetries */
} oracle_basis;

static int basis_number_bits(oracle_basis const *b) {
  return (b->particle_type == 1 ? 2 : 1) * b->number_sites;
}
/* Basis.hs:722-732 hasFixedHammingWeight */
static int basis_fixed_hamming(oracle_basis const *b) {
  if (b->particle_type == 0) return b->number_up != -1;
  return b->number_particles != -1;
}
static int basis_hamming_weight(oracle_basis const *b) {
  return b->particle_type == 0 ? b->number_up : b->number_particles;
}
static uint64_t low_ones(int h) { return h >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << h) - 1); }

/* Basis.hs:734-774 min/maxStateEstimate */
LS_ORACLE_API uint64_t oracle_min_state_estimate(oracle_basis const *b) {
  int const n = b->number_sites;
  if (b->particle_type == 1 && b->number_up != -1) {
    int const up = b->number_up, down = b->number_particles - b->number_up;
    return (low_ones(down) << n) | low_ones(up);
  }
  if (basis_fixed_hamming(b)) return low_ones(basis_hamming_weight(b));
  return 0;
}
LS_ORACLE_API uint64_t oracle_max_state_estimate(oracle_basis const *b) {
  int const n = b->number_sites;
  if (b->particle_type == 1 && b->number_up != -1) {
    int const up = b->number_up, down = b->number_particles - b->number_up;
    return ((low_ones(down) << (n - down)) << n) | (low_ones(up) << (n - up));
  }
  int const bits = basis_number_bits(b);
  if (basis_fixed_hamming(b)) {
    int const h = basis_hamming_weight(b);
    if (b->particle_type == 0 && b->spin_inversion != 0)
      return h == 0 ? 0 : low_ones(h) << (bits - h - 1);
    return low_ones(h) << (bits - h);
  }
  return low_ones(bits);
}

/* chapel/src/StatesEnumeration.chpl:242-346: emit, in ascending order, the
 * states of the basis.  Projected (:242-268): Gosper/+1 stepping over
 * [min,max], keep x iff is_representative flag && norm > 0.  Unprojected
 * (:269-289): everything in range, upper bound min(high, high^mask) under spin
 * inversion.  Spinful fermions with (n_up, n_down) (:290-326): product
 * enumeration, down sector is the slow index.  Returns malloc'ed array. */
LS_ORACLE_API uint64_t *oracle_enumerate_states(oracle_basis const *b,
                                                uint64_t *out_count) {
  oracle_vec out = {NULL, 0, 0};
  uint64_t lo = oracle_min_state_estimate(b);
  uint64_t hi = oracle_max_state_estimate(b);
  int const fixed = basis_fixed_hamming(b);
  if (b->particle_type == 1 && b->number_up != -1) {
    int const n = b->number_sites;
    uint64_t const mask = low_ones(n);
    uint64_t const minA = lo & mask, maxA = hi & mask;
    uint64_t const minB = (lo >> n) & mask, maxB = (hi >> n) & mask;
    for (uint64_t xb = minB;;) {
      for (uint64_t xa = minA;;) {
        vec_push(&out, (xb << n) | xa);
        if (xa == maxA) break;
        xa = next_state_fixed_hamming(xa);
      }
      if (xb == maxB) break;
      xb = next_state_fixed_hamming(xb);
    }
  } else if (b->particle_type == 0 && b->has_permutation_symmetries) {
    enum { BATCH = 10240 }; /* CommonParameters.chpl:5 */
    uint64_t *buf = (uint64_t *)malloc(BATCH * sizeof(uint64_t));
    uint8_t *flags = (uint8_t *)malloc(BATCH);
    double *norms = (double *)malloc(BATCH * sizeof(double));
    uint64_t x = lo;
    int done = 0;
    while (!done) {
      ptrdiff_t w = 0;
      for (;;) {
        buf[w++] = x;
        if (x == hi) { done = 1; break; }
        x = fixed ? next_state_fixed_hamming(x) : x + 1;
        if (w == BATCH) break;
      }
      oracle_is_representative(&b->group, w, buf, flags, norms);
      for (ptrdiff_t i = 0; i < w; ++i)
        if (flags[i] && norms[i] > 0) vec_push(&out, buf[i]);
    }
    free(buf); free(flags); free(norms);
  } else {
    if (b->particle_type == 0 && b->spin_inversion != 0) {
      uint64_t const mask = low_ones(b->number_sites);
      uint64_t const alt = hi ^ mask;
      if (alt < hi) hi = alt;
    }
    for (uint64_t x = lo;;) {
      vec_push(&out, x);
      if (x == hi) break;
      x = fixed ? next_state_fixed_hamming(x) : x + 1;
    }
  }
  *out_count = out.size;
  return out.data;
}

/* chapel/src/StatesEnumeration.chpl:242-268 on one chunk [lo, hi] (inclusive). */
static void enumerate_projected_chunk(oracle_basis const *b, uint64_t lo, uint64_t hi, oracle_vec *out) {
  enum { BATCH = 10240 }; /* CommonParameters.chpl:5 */
  int const fixed = basis_fixed_hamming(b);
  uint64_t *buf = (uint64_t *)malloc(BATCH * sizeof(uint64_t));
  uint8_t *flags = (uint8_t *)malloc(BATCH);
  double *norms = (double *)malloc(BATCH * sizeof(double));
  uint64_t x = lo;
  int done = 0;
  while (!done) {
    ptrdiff_t w = 0;
    for (;;) {
      buf[w++] = x;
      if (x == hi) { done = 1; break; }
      x = fixed ? next_state_fixed_hamming(x) : x + 1;
      if (w == BATCH) break;
    }
    oracle_is_representative(&b->group, w, buf, flags, norms);
    for (ptrdiff_t i = 0; i < w; ++i)
      if (flags[i] && norms[i] > 0) vec_push(out, buf[i]);
  }
  free(buf); free(flags); free(norms);
}

/* Projected enumeration of [lower, upper] (inclusive), split into chunks by
 * combinadic / linear index and run in parallel over the host cores like
 * StatesEnumeration.chpl:130-152 (determineEnumerationRanges) + :392-458
 * (forall over chunks); chunk outputs concatenate in order (:460-476). */
LS_ORACLE_API uint64_t *oracle_enumerate_range(oracle_basis const *b, uint64_t lower, uint64_t upper,
                                               uint64_t *out_count) {
  int const fixed = basis_fixed_hamming(b);
  int const hw = fixed ? basis_hamming_weight(b) : 0;
  init_binomials();
  uint64_t const first = fixed ? (uint64_t)oracle_fixed_hamming_state_to_index(lower) : lower;
  uint64_t const last = fixed ? (uint64_t)oracle_fixed_hamming_state_to_index(upper) : upper;
  uint64_t const total = last - first + 1;
  int nchunks = 8 * oracle_num_threads_impl();
  if ((uint64_t)nchunks > total) nchunks = (int)total;
  if (nchunks < 1) nchunks = 1;
  oracle_vec *parts = (oracle_vec *)calloc((size_t)nchunks, sizeof(oracle_vec));
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nchunks; ++c) {
    uint64_t const i0 = first + total / (uint64_t)nchunks * (uint64_t)c;
    uint64_t const i1 = (c == nchunks - 1) ? last : first + total / (uint64_t)nchunks * (uint64_t)(c + 1) - 1;
    uint64_t const lo = fixed ? oracle_fixed_hamming_index_to_state((int64_t)i0, hw) : i0;
    uint64_t const hi = fixed ? oracle_fixed_hamming_index_to_state((int64_t)i1, hw) : i1;
    enumerate_projected_chunk(b, lo, hi, &parts[c]);
  }
  size_t n = 0;
  for (int c = 0; c < nchunks; ++c) n += parts[c].size;
  uint64_t *out = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  size_t at = 0;
  for (int c = 0; c < nchunks; ++c) {
    if (parts[c].size) memcpy(out + at, parts[c].data, parts[c].size * sizeof(uint64_t));
    at += parts[c].size;
    free(parts[c].data);
  }
  free(parts);
  *out_count = n;
  return out;
}

LS_ORACLE_API void oracle_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* state_index: kernels/indexing.c:45-117 (bucket table) and :196-215, :273-325
 * (fixed-trip-count branchless search; -1 when absent). */
typedef struct oracle_index {
  ptrdiff_t number_states;
  uint64_t const *representatives;
  int prefix_bits;
  int shift;
  ptrdiff_t range_size;
  ptrdiff_t number_offsets;
  ptrdiff_t *offsets;
} oracle_index;

LS_ORACLE_API oracle_index *oracle_index_create(uint64_t const *reps,
                                                ptrdiff_t count,
                                                int number_bits,
                                                int prefix_bits) {
  oracle_index *c = (oracle_index *)calloc(1, sizeof(oracle_index));
  c->number_states = count;
  c->representatives = reps;
  c->prefix_bits = prefix_bits > number_bits ? number_bits : prefix_bits;
  c->shift = number_bits - c->prefix_bits;
  if (c->prefix_bits > 0) {
    ptrdiff_t const size = (ptrdiff_t)1 << c->prefix_bits;
    c->number_offsets = size + 1;
    c->offsets = (ptrdiff_t *)malloc((size_t)(size + 1) * sizeof(ptrdiff_t));
    ptrdiff_t first = 0;
    for (ptrdiff_t i = 0; i < size; ++i) {
      c->offsets[i] = first;
      while (first != count && (reps[first] >> c->shift) == (uint64_t)i) ++first;
    }
    c->offsets[size] = first;
    /* normalize_offset_ranges, indexing.c:74-92 */
    ptrdiff_t max_range = 0;
    for (ptrdiff_t i = 0; i < size; ++i) {
      ptrdiff_t const n = c->offsets[i + 1] - c->offsets[i];
      if (n > max_range) max_range = n;
    }
    for (ptrdiff_t i = 0; i < size; ++i)
      if (c->offsets[i] > count - max_range) c->offsets[i] = count - max_range;
    c->range_size = max_range;
  }
  return c;
}
LS_ORACLE_API void oracle_index_destroy(oracle_index *c) {
  if (c) { free(c->offsets); free(c); }
}
LS_ORACLE_API void oracle_state_index(oracle_index const *c, ptrdiff_t batch,
                                      uint64_t const *spins,
                                      ptrdiff_t *indices) {
#pragma omp parallel for schedule(static)
  for (ptrdiff_t b = 0; b < batch; ++b) {
    uint64_t const needle = spins[b];
    uint64_t const *base = c->representatives + c->offsets[needle >> c->shift];
    ptrdiff_t n = c->range_size;
    while (n > 1) { /* indexing.c:201-210 */
      ptrdiff_t const half = n / 2;
      n -= half;
      base = (base[half] < needle) ? base + half : base;
    }
    base += *base < needle;
    indices[b] = (*base == needle) ? base - c->representatives : -1;
  }
}

/* ------------------------------------------------------------------------- */
/* Operator terms: ls_hs_nonbranching_terms, lattice_symmetries_types.h:140-151
 * (single 64-bit word).  v is interleaved (re, im). */
typedef struct oracle_terms {
  int number_terms;
  double const *v; /* [T][2] */
  uint64_t const *m, *l, *r, *x, *s;
} oracle_terms;

/* kernels/reference.c:67-95 ls_internal_operator_apply_diag_x1 */
LS_ORACLE_API void oracle_apply_diag(oracle_terms const *t, ptrdiff_t batch,
                                     uint64_t const *alphas, double *ys,
                                     double const *xs) {
  if (t == NULL || t->number_terms == 0) {
    memset(ys, 0, (size_t)batch * sizeof(double));
    return;
  }
#pragma omp parallel for schedule(static)
  for (ptrdiff_t i = 0; i < batch; ++i) {
    double acc = 0;
    uint64_t const a = alphas[i];
    for (int k = 0; k < t->number_terms; ++k) {
      if ((a & t->m[k]) == t->r[k]) {
        int const sign = 1 - 2 * (__builtin_popcountll(a & t->s[k]) % 2);
        double const factor = (xs != NULL) ? sign * xs[i] : sign;
        acc += t->v[2 * k] * factor;
      }
    }
    ys[i] = acc;
  }
}

/* kernels/reference.c:97-134 ls_internal_operator_apply_off_diag_x1 for ONE
 * alpha; returns number of emitted (beta, coeff) pairs. */
static inline int apply_off_diag_one(oracle_terms const *t, uint64_t a,
                                     double x, uint64_t *betas,
                                     double *coeffs /* [.][2] */) {
  int n = 0;
  for (int k = 0; k < t->number_terms; ++k) {
    if ((a & t->m[k]) == t->r[k]) {
      int const sign = 1 - 2 * (__builtin_popcountll(a & t->s[k]) % 2);
      double const factor = sign * x;
      coeffs[2 * n + 0] = t->v[2 * k + 0] * factor;
      coeffs[2 * n + 1] = t->v[2 * k + 1] * factor;
      betas[n] = a ^ t->x[k];
      ++n;
    }
  }
  return n;
}

LS_ORACLE_API void oracle_apply_off_diag(oracle_terms const *t,
                                         ptrdiff_t batch,
                                         uint64_t const *alphas,
                                         uint64_t *betas, double *coeffs,
                                         ptrdiff_t *offsets, double const *xs) {
  offsets[0] = 0;
  ptrdiff_t off = 0;
  if (t != NULL)
    for (ptrdiff_t i = 0; i < batch; ++i) {
      off += apply_off_diag_one(t, alphas[i], xs ? xs[i] : 1.0, betas + off,
                                coeffs + 2 * off);
      offsets[i + 1] = off;
    }
  else
    for (ptrdiff_t i = 0; i < batch; ++i) offsets[i + 1] = 0;
}

/* Matvec, push form exactly as the reference assembles it:
 *   localMatrixVector       DistributedMatrixVector.chpl:1045-1058
 *   localDiagonal           :49-71   (y = diag * x; overwrites y)
 *   computeOffDiag dispatch BatchedOperator.chpl:264-282
 *     no projection         :139-164
 *     inversion only        :166-205  beta' = min(beta, beta^mask), c *= inv
 *     with projection       :207-253  state_info(betas ++ alphas),
 *                                     c = chi * (v*sign*x_i * n_beta / n_alpha)
 *   localProcess            :91-143  j = state_index(beta); y[j] += Re(c);
 *                                    c != 0 with j < 0 is an error.
 * y is always zero-initialised (the reference skips that when there are no
 * diagonal terms, :1054-1055, a latent bug noted in SURVEY 8a-6).
 * Rows [row_begin, row_end) are processed; returns the number of off-diagonal
 * matrix elements emitted, or -1 on an invalid index.  Threads use atomic adds
 * like ConcurrentAccessor.chpl:31-33. */
/* Optional hooks: the reference's OWN compiled kernels (oracle/_ref/libref.so = kernels/reference.c,
 * kernels/indexing.c) stand in for the two steps of the loop below that could be compiled from the reference tree:
 *   apply_off_diag : ls_internal_operator_apply_off_diag_x1(op, batch, alphas, betas, coeffs, offsets, xs)
 *   state_index    : ls_hs_state_index_binary_search_kernel(batch, spins, 1, indices, 1, data)
 * state_info stays the restatement (the reference's is Halide-generated).  Used by the CPU-baseline timing. */
typedef void (*ref_apply_off_diag_fn)(void const *op, ptrdiff_t batch, uint64_t const *alphas, uint64_t *betas,
                                      void *coeffs, ptrdiff_t *offsets, double const *xs);
typedef void (*ref_state_index_fn)(ptrdiff_t batch, uint64_t const *spins, ptrdiff_t spins_stride, ptrdiff_t *indices,
                                   ptrdiff_t indices_stride, void const *data);
static ref_apply_off_diag_fn g_ref_apply = NULL;
static ref_state_index_fn g_ref_index = NULL;
static void const *g_ref_operator = NULL;
static void const *g_ref_index_data = NULL;
LS_ORACLE_API void oracle_set_reference_hooks(void *apply_fn, void const *op, void *index_fn, void const *index_data) {
  g_ref_apply = (ref_apply_off_diag_fn)apply_fn;
  g_ref_operator = op;
  g_ref_index = (ref_state_index_fn)index_fn;
  g_ref_index_data = index_data;
}

/* block_stride: one block of 64 columns is processed every `block_stride` columns of [row_begin, row_end)
 * (64 = every column; larger = a uniform sample of the columns, for the CPU-baseline timing). */
LS_ORACLE_API int64_t oracle_matvec_strided(oracle_basis const *b,
                                    oracle_terms const *off,
                                    oracle_terms const *diag,
                                    oracle_index const *index,
                                    uint64_t const *reps, ptrdiff_t dim,
                                    ptrdiff_t row_begin, ptrdiff_t row_end, ptrdiff_t block_stride,
                                    double const *x, double *y,
                                    int zero_and_diag) {
  /* flags: bit 0 = zero y and apply the diagonal first (localDiagonal);
   * bit 1 = SAMPLING mode for the CPU-baseline timing only: `reps` is a sorted
   * PREFIX of the basis, matrix elements that leave it are searched for (same
   * work) and then dropped instead of raising the invalid-index error. */
  int const tolerate_missing = (zero_and_diag & 2) != 0;
  if (zero_and_diag & 1) {
    if (diag != NULL && diag->number_terms > 0)
      oracle_apply_diag(diag, dim, reps, y, x);
    else
      memset(y, 0, (size_t)dim * sizeof(double));
  }
  if (off == NULL || off->number_terms == 0) return 0;
  int const T = off->number_terms;
  int const with_projection = b->particle_type == 0 && b->has_permutation_symmetries;
  int const only_inversion =
      b->particle_type == 0 && !b->has_permutation_symmetries && b->spin_inversion != 0;
  int64_t total = 0;
  int bad = 0;
#pragma omp parallel reduction(+ : total)
  {
    enum { ROWS = 64 };
    size_t const cap = (size_t)ROWS * (size_t)(T + 1);
    uint64_t *tmp_spins = (uint64_t *)malloc(cap * sizeof(uint64_t));
    double *tmp_coeffs = (double *)malloc(cap * 2 * sizeof(double));
    uint64_t *betas = (uint64_t *)malloc(cap * sizeof(uint64_t));
    double *chars = (double *)malloc(cap * 2 * sizeof(double));
    double *norms = (double *)malloc(cap * sizeof(double));
    ptrdiff_t *idx = (ptrdiff_t *)malloc(cap * sizeof(ptrdiff_t));
    ptrdiff_t offsets[ROWS + 1];
    ptrdiff_t const stride = block_stride < ROWS ? ROWS : block_stride;
    ptrdiff_t const nblocks = row_end > row_begin ? (row_end - row_begin + stride - 1) / stride : 0;
#pragma omp for schedule(dynamic, 16)
    for (ptrdiff_t blk = 0; blk < nblocks; ++blk) {
      ptrdiff_t const r0 = row_begin + blk * stride;
      ptrdiff_t const count = (row_end - r0 < ROWS) ? row_end - r0 : ROWS;
      if (g_ref_apply != NULL)
        g_ref_apply(g_ref_operator, count, reps + r0, tmp_spins, tmp_coeffs, offsets, x + r0);
      else
        oracle_apply_off_diag(off, count, reps + r0, tmp_spins, tmp_coeffs,
                              offsets, x + r0);
      ptrdiff_t const n = offsets[count];
      total += n;
      double *cs;
      uint64_t *bs;
      if (with_projection) {
        memcpy(tmp_spins + n, reps + r0, (size_t)count * sizeof(uint64_t));
        /* single-threaded inner call: we are already inside a parallel region */
        oracle_state_info(&b->group, n + count, tmp_spins, betas, chars, norms);
        for (ptrdiff_t i = 0; i < count; ++i)
          for (ptrdiff_t k = offsets[i]; k < offsets[i + 1]; ++k) {
            /* cs[k] *= tempCoeffs[k] * norms[k] / norms[total + i] */
            double const tr = tmp_coeffs[2 * k] * norms[k] / norms[n + i];
            double const ti = tmp_coeffs[2 * k + 1] * norms[k] / norms[n + i];
            double const cr = chars[2 * k], ci = chars[2 * k + 1];
            chars[2 * k] = cr * tr - ci * ti;
            chars[2 * k + 1] = cr * ti + ci * tr;
          }
        cs = chars;
        bs = betas;
      } else if (only_inversion) {
        uint64_t const mask = low_ones(b->number_sites);
        for (ptrdiff_t k = 0; k < n; ++k) {
          uint64_t const inverted = tmp_spins[k] ^ mask;
          if (inverted < tmp_spins[k]) {
            tmp_spins[k] = inverted;
            tmp_coeffs[2 * k] *= b->spin_inversion;
            tmp_coeffs[2 * k + 1] *= b->spin_inversion;
          }
        }
        cs = tmp_coeffs;
        bs = tmp_spins;
      } else {
        cs = tmp_coeffs;
        bs = tmp_spins;
      }
      if (g_ref_index != NULL) g_ref_index(n, bs, 1, idx, 1, g_ref_index_data);
      else oracle_state_index(index, n, bs, idx);
      for (ptrdiff_t k = 0; k < n; ++k) {
        double const c = cs[2 * k]; /* coeffs[k]:real(64) */
        if (c != 0) {
          if (idx[k] >= 0) {
#pragma omp atomic
            y[idx[k]] += c;
          } else if (!tolerate_missing) {
            bad = 1;
          }
        }
      }
    }
    free(tmp_spins); free(tmp_coeffs); free(betas); free(chars); free(norms); free(idx);
  }
  return bad ? -1 : total;
}

LS_ORACLE_API int64_t oracle_matvec(oracle_basis const *b, oracle_terms const *off, oracle_terms const *diag,
                                    oracle_index const *index, uint64_t const *reps, ptrdiff_t dim,
                                    ptrdiff_t row_begin, ptrdiff_t row_end, double const *x, double *y,
                                    int zero_and_diag) {
  return oracle_matvec_strided(b, off, diag, index, reps, dim, row_begin, row_end, 64, x, y, zero_and_diag);
}

LS_ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
LS_ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
