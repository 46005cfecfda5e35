// matvec.cu -- sparse Hamiltonian matrix-vector product on the device.
//
// Replaces the Chapel driver chapel/src/DistributedMatrixVector.chpl
// (localDiagonal :49-71, localProcess :91-143, Producer.run :581-671,
// localMatrixVector :1045-1058, export :1090-1105), the batched operator
// chapel/src/BatchedOperator.chpl (:139-282 and the exports :298-357) and the
// scalar kernels kernels/reference.c:67-134.
//
// Design (not a port).  The reference pushes: for every column i it emits
// (beta, c) pairs, canonicalises beta, ranks it and does y[j] += c with atomic
// adds.  We pull: row i gathers its own contributions, so y is written once,
// coalesced, without atomics and in a deterministic order.  For the projected
// matrix  H~[j,i] = chi_g v_t sign_t(alpha_i) n_j / n_i  (BatchedOperator.chpl:
// 207-253) the pull form applies the ADJOINT terms to alpha_i:
//
//   y[i] = 1/n_i * sum_t' conj(chi_g v'_t sign'_t(alpha_i)) * n_j x[j]
//
//   T_t|a>   = v (-1)^{|a&s|} [a&m == r] |a^x>        (NonbranchingTerm.hs:24-33)
//   T_t^+|b> = conj(v) (-1)^{|x&s|} (-1)^{|b&s|} [b&m == l] |b^x>
//
// which is valid for any operator (Hermitian or not) that commutes with the
// symmetry group.  n_j is looked up (the norm is an orbit invariant, so the
// reference's state_info(beta).norm equals the stored norm of its
// representative) and folded into a pre-scaled copy xs[j] = n_j x[j].
//
// One fused kernel per matvec.  A block owns a tile of consecutive rows and
//   1. (thread per row) tests every adjoint term against alpha_i, evaluates
//      the diagonal and counts the matches; a block scan turns the counts
//      into positions in a shared-memory queue of (beta, term, sign);
//   2. (thread per 32 queue entries) transposes its 32 betas into bit planes
//      (bitslice.cuh) and walks the whole group: the image under element g is
//      a renaming of planes (plane_table.cuh), the running minimum costs two
//      LOP3 per plane per element for 32 states, and the index of the
//      minimising element is tracked in ten more planes;
//   3. the same thread transposes back, ranks its 32 representatives
//      (views.cuh state_index) and gathers conj(chi) w sign xs[j];
//   4. (thread per row) sums the row's segment of the queue in term order.
#include <cub/block/block_scan.cuh>

#include <algorithm>
#include <memory>

#include "bitslice.cuh"
#include "plane_table.cuh"
#include "state.hpp"

namespace lsb {

// ---- operator tables -----------------------------------------------------------
bool TermsDev::same_as(ls_hs_nonbranching_terms const *t) const {
  int const T = t == nullptr ? 0 : t->number_terms;
  if (T != number_terms) return false;
  if (T == 0) return true;
  size_t const n = (size_t)T;
  return memcmp(v.data(), t->v, n * 16) == 0 && memcmp(m.data(), t->m, n * 8) == 0 &&
         memcmp(l.data(), t->l, n * 8) == 0 && memcmp(r.data(), t->r, n * 8) == 0 &&
         memcmp(x.data(), t->x, n * 8) == 0 && memcmp(s.data(), t->s, n * 8) == 0;
}

void TermsDev::release() {
  cudaFree(d_v);
  cudaFree(d_m);
  cudaFree(d_l);
  cudaFree(d_r);
  cudaFree(d_x);
  cudaFree(d_s);
  d_v = nullptr;
  d_m = d_l = d_r = d_x = d_s = nullptr;
  number_terms = 0;
}

void TermsDev::upload(ls_hs_nonbranching_terms const *t) {
  release();
  int const T = t == nullptr ? 0 : t->number_terms;
  if (T == 0) return;
  LSB_CHECK(t->number_bits <= 64, "operators on more than 64 bits are not supported");
  size_t const n = (size_t)T;
  double const *tv = reinterpret_cast<double const *>(t->v);
  v.assign(tv, tv + 2 * n);
  m.assign(t->m, t->m + n);
  l.assign(t->l, t->l + n);
  r.assign(t->r, t->r + n);
  x.assign(t->x, t->x + n);
  s.assign(t->s, t->s + n);
  cudaStream_t st = runtime().stream;
  auto put = [&](auto **dst, void const *src, size_t bytes) {
    CUDA_CHECK(cudaMalloc(dst, bytes));
    CUDA_CHECK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st));
  };
  put(&d_v, v.data(), n * 16);
  put(&d_m, m.data(), n * 8);
  put(&d_l, l.data(), n * 8);
  put(&d_r, r.data(), n * 8);
  put(&d_x, x.data(), n * 8);
  put(&d_s, s.data(), n * 8);
  CUDA_CHECK(cudaStreamSynchronize(st));
  number_terms = T;
}

TermsView TermsDev::view() const { return TermsView{number_terms, d_v, d_m, d_l, d_r, d_x, d_s}; }

OperatorDev &operator_dev(ls_hs_operator const *op) {
  static std::unordered_map<ls_hs_operator const *, std::unique_ptr<OperatorDev>> cache;
  auto &slot = cache[op];
  if (!slot) slot = std::make_unique<OperatorDev>();
  OperatorDev &d = *slot;
  bool changed = false;
  if (!d.off.same_as(op->off_diag_terms)) {
    d.off.upload(op->off_diag_terms);
    std::vector<uint64_t> xs = d.off.x;
    std::sort(xs.begin(), xs.end());
    d.distinct_x = (int)(std::unique(xs.begin(), xs.end()) - xs.begin());
    changed = true;
  }
  if (!d.diag.same_as(op->diag_terms)) {
    d.diag.upload(op->diag_terms);
    changed = true;
  }
  if (changed) d.stats_rows = -1;
  return d;
}

// ---- kernel ----------------------------------------------------------------------
constexpr int kMvThreads = 128;
constexpr int kMvCap = kMvThreads * 32;  // queue entries per round
constexpr int kMvRowsPerThread = 2;
constexpr int kMvMaxRows = kMvThreads * kMvRowsPerThread;
constexpr int kMvIdxPlanes = 8;  // bit-sliced path: at most 256 distinct character values

enum : int { kModeNone = 0, kModeInversion = 1, kModeGroup = 2, kModeGroupScalar = 3 };

struct MatvecArgs {
  GroupView g;
  IndexView ix;
  TermsView off, diag;
  int mode;
  int number_idx_planes;  // ceil(log2(number of distinct characters))
  int debug_skip;         // LS_B200_MV_SKIP (profiling only): 1 = no orbit walk, 2 = no search/gather
  int number_chars;
  double2 const *cvals;  // distinct character values; chars[cidx] in the kernel
  int complex_vectors;  // x, xs, y hold interleaved (re, im)
  int rows_per_tile;
  int spin_inversion;
  uint64_t inversion_mask;
  int64_t row_begin, row_end;
  double const *norms;  // n_i of the representatives; nullptr when all 1
  double const *x;      // caller's vector (diagonal part)
  double const *xs;     // n_j x[j] (== x when norms is nullptr)
  double *y;            // rows [row_begin, row_end), i.e. y[0] is row_begin
  int *error_flag;
};

struct MvSmem {
  // all offsets in bytes from the dynamic shared memory base
  int off_m, off_l, off_x, off_s, off_w;      // off-diagonal terms
  int diag_m, diag_r, diag_s, diag_v;         // diagonal terms
  int chars;                                  // double2[|G|]
  int row_alpha, row_off, row_acc, row_diag;  // per-row arrays
  int queue, meta, cidx, values_im, planes;
  int total;
};

static MvSmem mv_layout(int T_off, int T_diag, int number_chars, int np, bool complex_vectors, bool group) {
  MvSmem L{};
  int p = 0;
  auto take = [&](int bytes) {
    int const at = p;
    p += (bytes + 15) & ~15;
    return at;
  };
  L.off_m = take(8 * T_off);
  L.off_l = take(8 * T_off);
  L.off_x = take(8 * T_off);
  L.off_s = take(8 * T_off);
  L.off_w = take(16 * T_off);
  L.diag_m = take(8 * T_diag);
  L.diag_r = take(8 * T_diag);
  L.diag_s = take(8 * T_diag);
  L.diag_v = take(16 * T_diag);
  L.chars = take(16 * number_chars);
  L.row_alpha = take(8 * kMvMaxRows);
  L.row_off = take(4 * (kMvMaxRows + 1));
  L.row_acc = take(16 * kMvMaxRows);
  L.row_diag = take(16 * kMvMaxRows);
  L.queue = take(8 * kMvCap);  // betas in, real parts of the contributions out
  L.meta = take(2 * kMvCap);
  L.cidx = take(kMvCap);
  L.values_im = take(complex_vectors ? 8 * kMvCap : 0);
  L.planes = take(group ? 4 * np * kMvThreads : 0);
  L.total = p;
  return L;
}

// Queue slot of entry q: lane k = q % 32 of word w = q / 32 lives at k * 128 + w,
// so that the thread owning word w touches bank w % 32 only (conflict-free).
__device__ __forceinline__ int queue_slot(int q) { return (q & 31) * kMvThreads + (q >> 5); }

// Stabiliser character sum of x read straight from the global tables; used on
// the (rare) path that decides whether a missing index is an error.
__device__ __noinline__ double stabiliser_sum_global(GroupView const &g, uint64_t x) {
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
__device__ __noinline__ void orbit_min_global(GroupView const &g, uint64_t x, uint64_t &rep, int &element, int &flipped) {
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

template <int NP, bool INV>
__global__ void __launch_bounds__(kMvThreads)
matvec_kernel(MatvecArgs const a, MvSmem const L) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t *t_m = reinterpret_cast<uint64_t *>(smem + L.off_m);
  uint64_t *t_l = reinterpret_cast<uint64_t *>(smem + L.off_l);
  uint64_t *t_x = reinterpret_cast<uint64_t *>(smem + L.off_x);
  uint64_t *t_s = reinterpret_cast<uint64_t *>(smem + L.off_s);
  double2 *t_w = reinterpret_cast<double2 *>(smem + L.off_w);
  uint64_t *d_m = reinterpret_cast<uint64_t *>(smem + L.diag_m);
  uint64_t *d_r = reinterpret_cast<uint64_t *>(smem + L.diag_r);
  uint64_t *d_s = reinterpret_cast<uint64_t *>(smem + L.diag_s);
  double2 *d_v = reinterpret_cast<double2 *>(smem + L.diag_v);
  double2 *chars = reinterpret_cast<double2 *>(smem + L.chars);
  uint64_t *row_alpha = reinterpret_cast<uint64_t *>(smem + L.row_alpha);
  uint32_t *row_off = reinterpret_cast<uint32_t *>(smem + L.row_off);
  double2 *row_acc = reinterpret_cast<double2 *>(smem + L.row_acc);
  double2 *row_diag = reinterpret_cast<double2 *>(smem + L.row_diag);
  uint64_t *queue = reinterpret_cast<uint64_t *>(smem + L.queue);
  double *values_re = reinterpret_cast<double *>(smem + L.queue);  // aliases the queue (same owner per slot)
  uint16_t *meta = reinterpret_cast<uint16_t *>(smem + L.meta);
  uint8_t *cidx = reinterpret_cast<uint8_t *>(smem + L.cidx);
  double *values_im = reinterpret_cast<double *>(smem + L.values_im);
  uint32_t *planes = reinterpret_cast<uint32_t *>(smem + L.planes);
  __shared__ uint32_t s_total;
  using BlockScan = cub::BlockScan<uint32_t, kMvThreads>;
  __shared__ typename BlockScan::TempStorage scan_storage;

  int const tid = threadIdx.x;
  int const T = a.off.number_terms;
  int const TD = a.diag.number_terms;
  int const G = a.g.number_masks;
  bool const cplx = a.complex_vectors != 0;

  // Stage the tables: adjoint off-diagonal terms (match on l, w = v (-1)^{|x&s|}).
  for (int t = tid; t < T; t += kMvThreads) {
    uint64_t const x = a.off.x[t], s = a.off.s[t];
    double2 v = a.off.v[t];
    if (__popcll(x & s) & 1) { v.x = -v.x; v.y = -v.y; }
    t_m[t] = a.off.m[t];
    t_l[t] = a.off.l[t];
    t_x[t] = x;
    t_s[t] = s;
    t_w[t] = v;
  }
  for (int t = tid; t < TD; t += kMvThreads) {
    d_m[t] = a.diag.m[t];
    d_r[t] = a.diag.r[t];
    d_s[t] = a.diag.s[t];
    d_v[t] = a.diag.v[t];
  }
  for (int j = tid; j < a.number_chars; j += kMvThreads) chars[j] = a.cvals[j];
  __syncthreads();

  int64_t const rows_total = a.row_end - a.row_begin;
  int const R = a.rows_per_tile;
  int64_t const tiles = (rows_total + R - 1) / R;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    int64_t const row0 = a.row_begin + tile * R;
    int const nrows = (int)min((int64_t)R, a.row_end - row0);

    // ---- phase 1: matches per row, diagonal ------------------------------------
    uint32_t counts[kMvRowsPerThread];
#pragma unroll
    for (int u = 0; u < kMvRowsPerThread; ++u) {
      int const r = tid * kMvRowsPerThread + u;
      uint32_t c = 0;
      if (r < nrows) {
        uint64_t const alpha = __ldg(a.ix.reps + row0 + r);
        row_alpha[r] = alpha;
        for (int t = 0; t < T; ++t) c += ((alpha & t_m[t]) == t_l[t]) ? 1u : 0u;
        double dr = 0.0, di = 0.0;
        for (int t = 0; t < TD; ++t)
          if ((alpha & d_m[t]) == d_r[t]) {
            double const sign = (__popcll(alpha & d_s[t]) & 1) ? -1.0 : 1.0;
            dr += sign * d_v[t].x;
            di += sign * d_v[t].y;
          }
        row_diag[r] = make_double2(dr, di);
        row_acc[r] = make_double2(0.0, 0.0);
      }
      counts[u] = c;
    }
    uint32_t offsets[kMvRowsPerThread];
    uint32_t total;
    BlockScan(scan_storage).ExclusiveSum(counts, offsets, total);
#pragma unroll
    for (int u = 0; u < kMvRowsPerThread; ++u) {
      int const r = tid * kMvRowsPerThread + u;
      if (r < nrows) row_off[r] = offsets[u];
    }
    if (tid == 0) {
      row_off[nrows] = total;
      s_total = total;
    }
    __syncthreads();
    total = s_total;

    for (uint32_t base = 0; base < total; base += kMvCap) {
      int const nwin = (int)min((uint32_t)kMvCap, total - base);

      // ---- phase 2: fill the queue window [base, base + nwin) -------------------
#pragma unroll
      for (int u = 0; u < kMvRowsPerThread; ++u) {
        int const r = tid * kMvRowsPerThread + u;
        if (r < nrows && row_off[r + 1] > base && row_off[r] < base + (uint32_t)nwin) {
          uint64_t const alpha = row_alpha[r];
          uint32_t q = row_off[r];
          for (int t = 0; t < T; ++t)
            if ((alpha & t_m[t]) == t_l[t]) {
              if (q >= base && q < base + (uint32_t)nwin) {
                int const slot = queue_slot((int)(q - base));
                queue[slot] = alpha ^ t_x[t];
                meta[slot] = (uint16_t)(t | ((__popcll(alpha & t_s[t]) & 1) << 15));
              }
              ++q;
            }
        }
      }
      __syncthreads();

      // ---- phase 3a: canonicalise (thread per 32 entries) ---------------------------
      // In:  queue[slot] = beta.  Out: queue[slot] = representative, cidx[slot] =
      // index (into chars[]) of the character of the minimising image.
      int const nwords = (nwin + 31) >> 5;
      // Whole warps enter (a warp-uniform condition) so that the group loop runs
      // converged and its per-element table reads use the uniform datapath; threads
      // past the last word carry lane 0 of word 0 and store nothing.
      if ((tid & ~31) < nwords) {
        int const lanes = max(0, min(32, nwin - 32 * tid));
        if (a.mode == kModeGroup) {
          uint32_t r[NP];
          {
            uint32_t lo[32], hi[32];
            uint64_t const pad = queue[lanes > 0 ? tid : 0];  // lane 0 of a live word is always valid
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              uint64_t const b = (k < lanes) ? queue[k * kMvThreads + tid] : pad;
              lo[k] = (uint32_t)b;
              hi[k] = (uint32_t)(b >> 32);
            }
            transpose32(lo);
            if (NP > 32) transpose32(hi);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              r[i] = (i < 32) ? lo[i] : hi[(i < 32) ? 0 : i - 32];
              planes[i * kMvThreads + tid] = r[i];
            }
          }
          __syncwarp();  // all 32 lanes are here (see above); a thread only reads back its own column
          unsigned char const *column = reinterpret_cast<unsigned char const *>(planes + tid);
          uint32_t idx[kMvIdxPlanes];
#pragma unroll
          for (int p = 0; p < kMvIdxPlanes; ++p) idx[p] = 0;  // character 0 == 1+0i: the input itself (generator.cpp:105-106)
          int const nbits = a.g.number_bits;
          int const nidx = a.number_idx_planes;
#pragma unroll 1
          for (int j = 0; j < ((a.debug_skip & 1) ? 0 : G); ++j) {
            uint16_t const *po = c_plane_offset + j * (NP + kPlaneRowExtra);
            // z = min(y, ~y) = y ^ top(y) when spin inversion is present (see basis_build.cu)
            uint32_t top = 0;
            if (INV) top = *reinterpret_cast<uint32_t const *>(column + po[NP]);
            uint32_t z[NP];
            uint32_t lt = 0;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              z[i] = *reinterpret_cast<uint32_t const *>(column + po[i]);
              if (INV) z[i] ^= (i < NP - 3 || i < nbits) ? top : 0u;  // padding planes stay zero
              lt = ((z[i] ^ r[i]) & r[i]) | (~(z[i] ^ r[i]) & lt);   // one LOP3: z < r, most significant plane last
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) r[i] = (lt & z[i]) | (~lt & r[i]);
            if (nidx > 0) {
              unsigned const ci = po[NP + 1];
              auto update = [&](int p) {
                uint32_t const c = 0u - ((ci >> p) & 1u);          // chi_j
                uint32_t const cf = 0u - ((ci >> (8 + p)) & 1u);   // inversion * chi_j
                uint32_t const nb = INV ? ((top & cf) | (~top & c)) : c;
                idx[p] = (lt & nb) | (~lt & idx[p]);
              };
              update(0);
              if (nidx > 1) update(1);
              if (nidx > 2) { update(2); update(3); }
              if (nidx > 4) { update(4); update(5); update(6); update(7); }
            }
          }
          // back to one state per word
          {
            uint32_t lo[32], hi[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              lo[i] = (i < NP) ? r[(i < NP) ? i : 0] : 0u;
              hi[i] = (i + 32 < NP) ? r[(i + 32 < NP) ? i + 32 : 0] : 0u;
            }
            transpose32(lo);
            if (NP > 32) transpose32(hi);
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < lanes) queue[k * kMvThreads + tid] = (NP > 32) ? (((uint64_t)hi[k] << 32) | lo[k]) : (uint64_t)lo[k];
          }
          {
            uint32_t info[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) info[i] = (i < kMvIdxPlanes) ? idx[(i < kMvIdxPlanes) ? i : 0] : 0u;
            if (nidx > 0) transpose32(info);
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < lanes) cidx[k * kMvThreads + tid] = (uint8_t)info[k];
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < lanes; ++k) {
            int const slot = k * kMvThreads + tid;
            uint64_t rep = queue[slot];
            unsigned c = 0;
            if (a.mode == kModeGroupScalar) {
              int e, flipped;
              uint64_t r2;
              orbit_min_global(a.g, rep, r2, e, flipped);
              rep = r2;
              if (e >= 0) {
                unsigned const ci = __ldg(a.g.cinfo + e);
                c = flipped ? (ci >> 8) : (ci & 0xffu);
              }
            } else if (a.mode == kModeInversion) {
              // BatchedOperator.chpl:187-199
              uint64_t const inverted = rep ^ a.inversion_mask;
              if (inverted < rep) {
                rep = inverted;
                c = 1;
              }
            }
            queue[slot] = rep;
            cidx[slot] = (uint8_t)c;
          }
        }
      }
      __syncthreads();

      // ---- phase 3b: rank + gather, eight independent searches in flight per thread ----
      if (tid < nwords && !(a.debug_skip & 2)) {
        int const lanes = min(32, nwin - 32 * tid);
        constexpr int B = 8;
#pragma unroll 1
        for (int k0 = 0; k0 < 32; k0 += B) {
          uint64_t needle[B];
          int64_t lo[B], n[B];
#pragma unroll
          for (int u = 0; u < B; ++u) {
            bool const live = k0 + u < lanes;
            needle[u] = live ? queue[(k0 + u) * kMvThreads + tid] : 0;
            index_window(a.ix, needle[u], live, lo[u], n[u]);
          }
          if (!a.ix.identity) {
#pragma unroll 1
            for (int s = 0; s < a.ix.steps; ++s) {
#pragma unroll
              for (int u = 0; u < B; ++u) {
                int64_t const half = n[u] >> 1;
                int64_t const mid = lo[u] + half;
                uint64_t const v = n[u] > 0 ? __ldg(a.ix.reps + mid) : ~uint64_t(0);
                bool const less = v < needle[u];
                lo[u] = less ? mid + 1 : lo[u];
                n[u] = less ? n[u] - half - 1 : half;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < B; ++u) {
            int const slot = (k0 + u) * kMvThreads + tid;
            bool const live = k0 + u < lanes;
            int64_t j = -1;
            if (live) {
              if (a.ix.identity) j = (int64_t)needle[u];
              else if (lo[u] < a.ix.number_states && __ldg(a.ix.reps + lo[u]) == needle[u]) j = lo[u];
            }
            double vr = 0.0, vi = 0.0;
            if (live) {
              unsigned const mt = meta[slot];
              double2 w = t_w[mt & 0x7fffu];
              if (mt & 0x8000u) { w.x = -w.x; w.y = -w.y; }
              if (j >= 0) {
                double2 const c = chars[cidx[slot]];
                // conj(chi) * w * xs[j]
                double const fr = c.x * w.x + c.y * w.y;
                double const fi = c.x * w.y - c.y * w.x;
                if (cplx) {
                  double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.xs) + j);
                  vr = fr * xv.x - fi * xv.y;
                  vi = fr * xv.y + fi * xv.x;
                } else {
                  vr = fr * __ldg(a.xs + j);
                }
              } else if (w.x != 0.0 || w.y != 0.0) {
                // Not in the basis: fine when its norm vanishes (the reference
                // multiplies by n_beta = 0), an error otherwise
                // (DistributedMatrixVector.chpl:127-135).
                bool bad = true;
                if (a.mode == kModeGroup || a.mode == kModeGroupScalar)
                  bad = stabiliser_sum_global(a.g, needle[u]) > kNormThreshold;
                if (bad) atomicOr(a.error_flag, 1);
              }
            }
            values_re[slot] = vr;
            if (cplx) values_im[slot] = vi;
          }
        }
      }
      __syncthreads();

      // ---- phase 4: per-row sums in term order -----------------------------------
#pragma unroll
      for (int u = 0; u < kMvRowsPerThread; ++u) {
        int const r = tid * kMvRowsPerThread + u;
        if (r < nrows) {
          uint32_t const q0 = max(row_off[r], base);
          uint32_t const q1 = min(row_off[r + 1], base + (uint32_t)nwin);
          if (q0 < q1) {
            double sr = row_acc[r].x, si = row_acc[r].y;
            for (uint32_t q = q0; q < q1; ++q) {
              int const slot = queue_slot((int)(q - base));
              sr += values_re[slot];
              if (cplx) si += values_im[slot];
            }
            row_acc[r] = make_double2(sr, si);
          }
        }
      }
      __syncthreads();
    }

    // ---- write y -------------------------------------------------------------------
    for (int r = tid; r < nrows; r += kMvThreads) {
      int64_t const row = row0 + r;
      double const ni = a.norms != nullptr ? __ldg(a.norms + row) : 1.0;
      double2 const acc = row_acc[r];
      double2 const dg = row_diag[r];
      int64_t const out = row - a.row_begin;
      if (cplx) {
        double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.x) + row);
        double2 res;
        res.x = acc.x / ni + (dg.x * xv.x - dg.y * xv.y);
        res.y = acc.y / ni + (dg.x * xv.y + dg.y * xv.x);
        reinterpret_cast<double2 *>(a.y)[out] = res;
      } else {
        // kernels/reference.c:84-91 uses creal(v) only
        a.y[out] = acc.x / ni + dg.x * __ldg(a.x + row);
      }
    }
    __syncthreads();
  }
}

// xs[j] = n_j x[j]
__global__ void __launch_bounds__(256)
prescale_kernel(int64_t n, int complex_vectors, double const *__restrict__ norms,
                double const *__restrict__ x, double *__restrict__ xs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double const s = norms[i];
    if (complex_vectors) {
      double2 const v = reinterpret_cast<double2 const *>(x)[i];
      reinterpret_cast<double2 *>(xs)[i] = make_double2(s * v.x, s * v.y);
    } else {
      xs[i] = s * x[i];
    }
  }
}

// Number of off-diagonal matrix elements in a row range = matching (row, term)
// pairs of the adjoint list (equal, summed over all rows of a symmetric
// operator, to the (alpha, term) pairs the reference's push form emits,
// kernels/reference.c:109-129).
__global__ void __launch_bounds__(256)
count_elements_kernel(TermsView off, uint64_t const *__restrict__ reps, int64_t row_begin, int64_t row_end,
                      unsigned long long *__restrict__ out) {
  unsigned long long local = 0;
  for (int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_end;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const alpha = reps[i];
    for (int t = 0; t < off.number_terms; ++t) local += ((alpha & __ldg(off.m + t)) == __ldg(off.l + t)) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(out, local);
}

void ensure_norms(IndexData &ix, GroupData const &g);

using MatvecKernel = void (*)(MatvecArgs, MvSmem);
template <int NP>
static MatvecKernel pick_mv_inv(bool inv) {
  return inv ? matvec_kernel<NP, true> : matvec_kernel<NP, false>;
}
static MatvecKernel pick_matvec_kernel(int np, bool inv) {
  switch (np) {
    case 4: return pick_mv_inv<4>(inv);
    case 8: return pick_mv_inv<8>(inv);
    case 12: return pick_mv_inv<12>(inv);
    case 16: return pick_mv_inv<16>(inv);
    case 20: return pick_mv_inv<20>(inv);
    case 24: return pick_mv_inv<24>(inv);
    case 28: return pick_mv_inv<28>(inv);
    case 32: return pick_mv_inv<32>(inv);
    case 36: return pick_mv_inv<36>(inv);
    case 40: return pick_mv_inv<40>(inv);
    case 44: return pick_mv_inv<44>(inv);
    case 48: return pick_mv_inv<48>(inv);
    case 52: return pick_mv_inv<52>(inv);
    case 56: return pick_mv_inv<56>(inv);
    case 60: return pick_mv_inv<60>(inv);
    case 64: return pick_mv_inv<64>(inv);
  }
  return nullptr;
}

static int64_t count_elements(OperatorDev &od, IndexData const &ix, int64_t row_begin, int64_t row_end) {
  Runtime &rt = runtime();
  if (od.off.number_terms == 0 || row_end <= row_begin) return 0;
  static unsigned long long *d_counter = nullptr;
  if (d_counter == nullptr) CUDA_CHECK(cudaMalloc(&d_counter, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), rt.stream));
  int64_t const n = row_end - row_begin;
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
  count_elements_kernel<<<blocks, 256, 0, rt.stream>>>(od.off.view(), ix.d_reps, row_begin, row_end, d_counter);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, d_counter, sizeof h, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  return (int64_t)h;
}

struct MatvecScratch {
  DeviceBuffer<double> x, xs, y;
  int *d_error = nullptr;
  double2 *d_plain_chars = nullptr;  // {1, +1, -1}: character table of the unprojected / inversion-only modes
};
static MatvecScratch &mv_scratch() {
  static MatvecScratch s;
  return s;
}

// y[row_begin:row_end] = (H x)[row_begin:row_end]; x, y in device memory.
static void matvec_device(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, double const *d_x,
                          double *d_y, bool complex_vectors) {
  Runtime &rt = runtime();
  ls_hs_basis const *basis = op->basis;
  IndexData *ix = index_of(basis);
  LSB_CHECK(ix != nullptr, "basis is not built: call ls_hs_basis_build / ls_hs_build_representatives first");
  int64_t const dim = ix->number_states;
  LSB_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= dim, "invalid row range");
  if (row_begin == row_end) return;
  OperatorDev &od = operator_dev(op);
  BasisInfo const info = basis_info(basis);
  MatvecScratch &sc = mv_scratch();
  if (sc.d_error == nullptr) {
    CUDA_CHECK(cudaMalloc(&sc.d_error, sizeof(int)));
    CUDA_CHECK(cudaMemsetAsync(sc.d_error, 0, sizeof(int), rt.stream));
    double const plain[6] = {1.0, 0.0, 1.0, 0.0, -1.0, 0.0};
    CUDA_CHECK(cudaMalloc(&sc.d_plain_chars, sizeof plain));
    CUDA_CHECK(cudaMemcpy(sc.d_plain_chars, plain, sizeof plain, cudaMemcpyHostToDevice));
  }

  MatvecArgs a{};
  a.ix = ix->view();
  a.off = od.off.view();
  a.diag = od.diag.view();
  a.complex_vectors = complex_vectors ? 1 : 0;
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.x = d_x;
  a.xs = d_x;
  a.y = d_y;
  a.error_flag = sc.d_error;
  a.spin_inversion = basis->spin_inversion;
  a.inversion_mask = basis->number_sites >= 64 ? ~uint64_t(0) : ((uint64_t(1) << basis->number_sites) - 1);
  a.mode = kModeNone;
  if (char const *dbg = getenv("LS_B200_MV_SKIP")) a.debug_skip = atoi(dbg);
  a.cvals = sc.d_plain_chars;
  a.number_chars = 1;
  int np = 4;
  bool inv = false;
  bool bitsliced = false;
  if (info.has_permutation_symmetries) {
    GroupData const &g = *info.group;
    ensure_norms(*ix, g);
    a.g = g.view();
    a.norms = ix->d_norms;
    np = std::max(4, (g.number_bits + 3) / 4 * 4);
    inv = g.spin_inversion != 0;
    char const *env = getenv("LS_B200_MATVEC");
    bool const want_scalar = env != nullptr && strcmp(env, "scalar") == 0;
    LSB_CHECK(!g.cinfo.empty(), "symmetry sectors with more than 256 distinct character values are not supported");
    a.cvals = g.d_cvals;
    a.number_chars = (int)(g.cvals.size() / 2);
    while ((1 << a.number_idx_planes) < a.number_chars) ++a.number_idx_planes;
    bitsliced = !want_scalar && upload_plane_offsets(g, np, kMvThreads);
    a.mode = bitsliced ? kModeGroup : kModeGroupScalar;
    // pre-scaled copy of x
    size_t const words = (size_t)dim * (complex_vectors ? 2 : 1);
    double *xs = sc.xs.reserve(words);
    unsigned const blocks = (unsigned)std::min<int64_t>((dim + 255) / 256, (int64_t)rt.sm_count * 16);
    prescale_kernel<<<blocks, 256, 0, rt.stream>>>(dim, a.complex_vectors, ix->d_norms, d_x, xs);
    count_launch();
    CUDA_CHECK(cudaGetLastError());
    a.xs = xs;
  } else if (info.has_spin_inversion) {
    a.mode = kModeInversion;
    // cidx 1 = the spin-inversion character: entry 1 of the plain table is +1, entry 2 is -1
    a.cvals = sc.d_plain_chars + (basis->spin_inversion < 0 ? 1 : 0);
    a.number_chars = 2;
  }

  // tile size from the average number of matrix elements per row
  if (od.stats_index != (void const *)ix || od.stats_rows != dim) {
    od.stats_elements = count_elements(od, *ix, 0, dim);
    od.stats_index = ix;
    od.stats_rows = dim;
  }
  double const avg = dim > 0 ? (double)od.stats_elements / (double)dim : 0.0;
  int rows = kMvMaxRows;
  if (avg > 0.0) rows = (int)std::min<double>(kMvMaxRows, std::max(1.0, 0.94 * kMvCap / avg));
  a.rows_per_tile = rows;

  bool const group_planes = a.mode == kModeGroup;
  MvSmem const L = mv_layout(a.off.number_terms, a.diag.number_terms, a.number_chars, np, complex_vectors, group_planes);
  LSB_CHECK((size_t)L.total <= rt.smem_optin, "operator / symmetry tables do not fit in shared memory");
  LSB_CHECK(a.off.number_terms < 0x8000, "too many off-diagonal terms");
  MatvecKernel kernel = group_planes ? pick_matvec_kernel(np, inv) : matvec_kernel<4, false>;
  LSB_CHECK(kernel != nullptr, "unsupported number of bits");
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  int per_sm = 1;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMvThreads, (size_t)L.total));
  per_sm = std::max(per_sm, 1);
  int64_t const tiles = (row_end - row_begin + rows - 1) / rows;
  unsigned const blocks = (unsigned)std::min<int64_t>(tiles, (int64_t)rt.sm_count * per_sm);
  CUDA_CHECK(cudaEventRecord(rt.ev0, rt.stream));
  kernel<<<blocks, kMvThreads, (size_t)L.total, rt.stream>>>(a, L);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(rt.ev1, rt.stream));
}

// Blocks until the stream drains, then reports kernel time and sector errors.
static bool matvec_finish() {
  Runtime &rt = runtime();
  MatvecScratch &sc = mv_scratch();
  int flag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&flag, sc.d_error, sizeof(int), cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, rt.ev0, rt.ev1) == cudaSuccess) rt.last_matvec_ms = ms;
  if (flag != 0) {
    CUDA_CHECK(cudaMemsetAsync(sc.d_error, 0, sizeof(int), rt.stream));
    return false;
  }
  return true;
}

static char const *kInvalidIndexMessage =
    "matrix_vector_product: the operator maps a basis state outside of the basis with a non-zero "
    "coefficient (invalid index); it does not respect the symmetries of the basis";

}  // namespace lsb

using namespace lsb;

extern "C" {

// chapel/src/DistributedMatrixVector.chpl:1090-1105 (host pointers, float64 only)
void ls_chpl_matrix_vector_product(ls_hs_operator *op, int num_vectors, double const *x, double *y) {
  int const number_bits = (op->basis->particle_type == LS_HS_SPINFUL_FERMION ? 2 : 1) * op->basis->number_sites;
  if (number_bits > 64) {
    ls_hs_error("bases with more than 64 bits are not yet implemented");
    return;
  }
  if (num_vectors != 1) {
    ls_hs_error("applying the Operator to more than 1 vector is not yet implemented");
    return;
  }
  bool ok = true;
  guarded(__func__, [&] {
    IndexData *ix = index_of(op->basis);
    LSB_CHECK(ix != nullptr, "basis is not built");
    int64_t const dim = ix->number_states;
    if (dim == 0) return;
    MatvecScratch &sc = mv_scratch();
    cudaStream_t s = runtime().stream;
    double *d_x = sc.x.reserve((size_t)dim);
    double *d_y = sc.y.reserve((size_t)dim);
    CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)dim, cudaMemcpyHostToDevice, s));
    matvec_device(op, 0, dim, d_x, d_y, false);
    CUDA_CHECK(cudaMemcpyAsync(y, d_y, sizeof(double) * (size_t)dim, cudaMemcpyDeviceToHost, s));
    ok = matvec_finish();
  });
  if (!ok) ls_hs_error(kInvalidIndexMessage);
}

int ls_b200_matvec_device(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, double const *x_dev,
                          double *y_dev) {
  int status = -1;
  guarded(__func__, [&] {
    matvec_device(op, row_begin, row_end, x_dev, y_dev, false);
    status = 0;
  });
  return status;
}

int ls_b200_matvec_device_c128(ls_hs_operator const *op, int64_t row_begin, int64_t row_end,
                               ls_hs_scalar const *x_dev, ls_hs_scalar *y_dev) {
  int status = -1;
  guarded(__func__, [&] {
    matvec_device(op, row_begin, row_end, reinterpret_cast<double const *>(x_dev),
                  reinterpret_cast<double *>(y_dev), true);
    status = 0;
  });
  return status;
}

int ls_b200_matvec_sync(void) {
  int status = -1;
  bool ok = true;
  guarded(__func__, [&] {
    if (mv_scratch().d_error == nullptr) {
      CUDA_CHECK(cudaStreamSynchronize(runtime().stream));
    } else {
      ok = matvec_finish();
    }
    status = 0;
  });
  if (!ok) {
    ls_hs_error(kInvalidIndexMessage);
    return 1;
  }
  return status;
}

int64_t ls_b200_count_matrix_elements(ls_hs_operator const *op, int64_t row_begin, int64_t row_end) {
  int64_t n = -1;
  guarded(__func__, [&] {
    IndexData *ix = index_of(op->basis);
    LSB_CHECK(ix != nullptr, "basis is not built");
    n = count_elements(operator_dev(op), *ix, row_begin, row_end);
  });
  return n;
}

}  // extern "C"
