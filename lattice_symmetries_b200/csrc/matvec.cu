// matvec.cu -- sparse Hamiltonian matrix-vector product on the device.
//
// Replaces the Chapel driver chapel/src/DistributedMatrixVector.chpl
// (localDiagonal :49-71, localProcess :91-143, Producer.run :581-671,
// localMatrixVector :1045-1058, export :1090-1105), the batched operator
// chapel/src/BatchedOperator.chpl (:139-282) and the scalar kernels
// kernels/reference.c:67-134.
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
// The rows are processed in chunks whose intermediates stay L2-resident; per
// chunk three kernels run back to back, each at its own best occupancy:
//   row_count_kernel   thread per row: number of matching adjoint terms; a
//                      device scan turns the counts into CSR-style offsets;
//   orbit_kernel       thread per 32 matrix elements (integer-issue bound):
//                      regenerates its 32 betas from the offsets, transposes
//                      them into bit planes (bitslice.cuh) and walks the whole
//                      group -- the image under g is a renaming of planes
//                      (plane_table.cuh), the running minimum costs three LOP3
//                      per plane per element for 32 states -- and writes the
//                      representative + the index of the minimising character;
//   gather_kernel      thread per row (latency / HBM bound): ranks each
//                      representative (bucket table + branchless search over
//                      the sorted representatives, several searches in flight),
//                      gathers conj(chi) w sign xs[j], sums in term order and
//                      writes y[i] once.
// Unprojected bases (and spin-inversion-only ones) skip the first two kernels.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <memory>
#include <type_traits>

#include "matvec_args.cuh"


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

static std::unordered_map<ls_hs_operator const *, std::unique_ptr<OperatorDev>> &operator_cache() {
  static std::unordered_map<ls_hs_operator const *, std::unique_ptr<OperatorDev>> cache;
  return cache;
}

OperatorDev &operator_dev(ls_hs_operator const *op) {
  auto &cache = operator_cache();
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
  if (changed) {
    d.stats_rows = -1;
    ++d.version;
  }
  return d;
}

// counts[r] = number of adjoint terms matching row chunk_begin + r; counts[chunk_rows] = 0.
__global__ void __launch_bounds__(256)
row_count_kernel(MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  AdjointTerms terms;
  terms.stage(smem, a.off, false);
  __syncthreads();
  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > a.chunk_rows) return;
  uint32_t c = 0;
  if (r < a.chunk_rows) {
    uint64_t const alpha = __ldg(a.rows + a.chunk_begin + r);
    for (int t = 0; t < terms.T; ++t) c += ((alpha & terms.m[t]) == terms.l[t]) ? 1u : 0u;
  }
  a.counts[r] = c;
}

// Thread per row: sum the row's values in term order, add the diagonal, write y once.
template <bool CPLX>
__global__ void __launch_bounds__(256)
row_sum_kernel(MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  int const TD = a.diag.number_terms;
  double2 *d_v = reinterpret_cast<double2 *>(smem);
  uint64_t *d_m = reinterpret_cast<uint64_t *>(d_v + TD);
  uint64_t *d_r = d_m + TD;
  uint64_t *d_s = d_r + TD;
  for (int t = threadIdx.x; t < TD; t += blockDim.x) {
    d_m[t] = a.diag.m[t];
    d_r[t] = a.diag.r[t];
    d_s[t] = a.diag.s[t];
    d_v[t] = a.diag.v[t];
  }
  __syncthreads();
  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.chunk_rows) return;
  int64_t const vec = blockIdx.y;  // block matvec: one grid row per vector
  int64_t const row = a.chunk_begin + r;
  uint64_t const alpha = __ldg(a.rows + row);
  uint32_t const qa = __ldg(a.offsets + r), qb = __ldg(a.offsets + r + 1);
  double acc_r = 0.0, acc_i = 0.0;
  for (uint32_t q = qa; q < qb; ++q) {
    if (CPLX) {
      double2 const v = __ldcs(reinterpret_cast<double2 const *>(a.vals) + vec * a.vals_stride + q);
      acc_r += v.x;
      acc_i += v.y;
    } else {
      acc_r += __ldcs(a.vals + vec * a.vals_stride + q);
    }
  }
  double dr = 0.0, di = 0.0;
  for (int k = 0; k < TD; ++k)
    if ((alpha & d_m[k]) == d_r[k]) {
      double const sign = (__popcll(alpha & d_s[k]) & 1) ? -1.0 : 1.0;
      dr += sign * d_v[k].x;
      di += sign * d_v[k].y;
    }
  double const ni = a.norms != nullptr ? __ldg(a.norms + row) : 1.0;
  int64_t const out = row - a.row_begin;
  if (CPLX) {
    double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.x) + vec * a.x_stride + row);
    double2 res;
    res.x = acc_r / ni + (dr * xv.x - di * xv.y);
    res.y = acc_i / ni + (dr * xv.y + di * xv.x);
    reinterpret_cast<double2 *>(a.y)[vec * a.y_stride + out] = res;
  } else {
    a.y[vec * a.y_stride + out] = acc_r / ni + dr * __ldg(a.x + vec * a.x_stride + row);  // kernels/reference.c:84-91 uses creal(v) only
  }
}

// ---- split path: orbit_kernel -> rank_gather_kernel -> row_combine_kernel ------------
// The ranking + gather is pure memory latency; on its own it runs at full occupancy
// (thread per matrix element, kRankBatch independent searches per thread, few registers,
// no shared memory) instead of sharing the register-heavy orbit kernel's 16-20 warps.
//   gathered[q] = n_j x_j of the representative of matrix element q (0 when the state has
//   no index; kMissBits when it has none although its norm is positive -- an error unless
//   the term's coefficient vanishes, which row_combine_kernel decides).
#ifndef LS_RANK_MINBLOCKS
#define LS_RANK_MINBLOCKS 5  // <= 48 registers
#endif
constexpr int kRankThreads = 256;
#ifndef LS_RANK_BATCH
#define LS_RANK_BATCH 4
#endif
constexpr int kRankBatch = LS_RANK_BATCH;
constexpr unsigned long long kMissBits = 0x7ff8dead00000001ull;  // a quiet NaN no computation produces

template <class Low, bool Wide = false, bool Sorted = false>
__global__ void __launch_bounds__(kRankThreads, LS_RANK_MINBLOCKS)  // register cap, also for the (rare) noinline norm check
rank_gather_kernel(__grid_constant__ MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  // product mode (q_tsign recorded): the values leave this kernel already multiplied by
  // conj(chi) w sign, and a plain per-row sum finishes the job
  bool const product = a.q_tsign != nullptr;
  double2 *tw = reinterpret_cast<double2 *>(smem);
  double2 *chars = tw + a.off.number_terms;
  if (product) {
    for (int t = threadIdx.x; t < a.off.number_terms; t += blockDim.x) {
      double2 v = a.off.v[t];
      if (__popcll(a.off.x[t] & a.off.s[t]) & 1) { v.x = -v.x; v.y = -v.y; }
      tw[t] = v;
    }
    for (int j = threadIdx.x; j < a.number_chars; j += blockDim.x) chars[j] = a.cvals[j];
    __syncthreads();
  }
  uint32_t const total = a.offsets[a.chunk_rows];
  int const lane = threadIdx.x & 31;
  bool const cplx = a.complex_vectors != 0;
  // persistent: the grid is sized to the machine, a warp strides over tiles of 32 * kRankBatch elements
  uint64_t const warp_stride = (uint64_t)gridDim.x * (kRankThreads / 32) * (32 * kRankBatch);
  for (uint64_t warp_q0 = ((uint64_t)blockIdx.x * (kRankThreads / 32) + (threadIdx.x >> 5)) * (32 * kRankBatch);
       warp_q0 < total; warp_q0 += warp_stride) {
  uint64_t needle[kRankBatch];
  bool live[kRankBatch];
#pragma unroll
  for (int u = 0; u < kRankBatch; ++u) {
    uint64_t const q = warp_q0 + (uint64_t)u * 32 + lane;
    live[u] = q < total;
    needle[u] = live[u] ? (Sorted ? __ldcs(a.q_sorted + q) : __ldcs(a.q_rep + q)) : 0;
  }
  int64_t j[kRankBatch];
  if (a.debug_skip & 8) {  // profiling only: no index search
#pragma unroll
    for (int u = 0; u < kRankBatch; ++u) j[u] = (int64_t)(needle[u] % (uint64_t)a.ix.number_states);
  } else if constexpr (std::is_void<Low>::value) index_find<kRankBatch>(a.ix, needle, live, j);
  else index_find32<Low, kRankBatch, Wide>(a.ix, needle, live, j);
  if (a.debug_skip & 4) {  // profiling only: no random gather
#pragma unroll
    for (int u = 0; u < kRankBatch; ++u) j[u] = j[u] >= 0 ? (j[u] & 1023) : j[u];
  }
  double2 xv[kRankBatch];
#pragma unroll
  for (int u = 0; u < kRankBatch; ++u) {
    xv[u] = make_double2(0.0, 0.0);
    if (j[u] >= 0) {
      if (cplx) xv[u] = __ldg(reinterpret_cast<double2 const *>(a.xs) + j[u]);
      else xv[u].x = __ldg(a.xs + j[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < kRankBatch; ++u) {
    if (!live[u]) continue;
    uint64_t q = warp_q0 + (uint64_t)u * 32 + lane;
    if (Sorted) q = __ldcs(a.perm + q);  // sorted ranking: back to the element's CSR position
    bool const missing = j[u] < 0;
    double fr = 1.0, fi = 0.0;
    if (product) {
      unsigned const ts = __ldcs(a.q_tsign + q);
      double2 w = tw[ts & 0x7fffu];
      if (ts & 0x8000u) { w.x = -w.x; w.y = -w.y; }
      double2 const ch = chars[a.number_idx_planes > 0 ? __ldcs(a.q_cidx + q) : 0];
      fr = ch.x * w.x + ch.y * w.y;  // conj(chi) * w
      fi = ch.x * w.y - ch.y * w.x;
      // not in the basis: fine when its norm vanishes, an error otherwise (DistributedMatrixVector.chpl:127-135)
      if (missing && (fr != 0.0 || fi != 0.0) && stabiliser_sum_global(a.g, needle[u]) > kNormThreshold)
        atomicOr(a.error_flag, 1);
      xv[u] = make_double2(fr * xv[u].x - fi * xv[u].y, fr * xv[u].y + fi * xv[u].x);
    } else if (missing && stabiliser_sum_global(a.g, needle[u]) > kNormThreshold) {
      xv[u].x = __longlong_as_double((long long)kMissBits);
    }
    if (cplx) reinterpret_cast<double2 *>(a.vals)[q] = xv[u];
    else a.vals[q] = xv[u].x;
    // block matvec: the other vectors reuse the rank and the coefficient (product mode only)
    for (int v = 1; v < a.number_vectors; ++v) {
      double2 o = make_double2(0.0, 0.0);
      if (!missing) {
        if (cplx) {
          double2 const xo = __ldg(reinterpret_cast<double2 const *>(a.xs) + v * a.xs_stride + j[u]);
          o = make_double2(fr * xo.x - fi * xo.y, fr * xo.y + fi * xo.x);
        } else {
          o.x = fr * __ldg(a.xs + v * a.xs_stride + j[u]);
        }
      }
      if (cplx) reinterpret_cast<double2 *>(a.vals)[v * a.vals_stride + q] = o;
      else a.vals[v * a.vals_stride + q] = o.x;
    }
  }
  }  // tiles of this warp
}

// Thread per row: conj(chi) w sign times the gathered values in term order, the
// diagonal, one write of y.
template <bool CPLX>
__global__ void __launch_bounds__(kGatherThreads)
row_combine_kernel(__grid_constant__ MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  AdjointTerms terms;
  terms.stage(smem, a.off, true);
  size_t p = (AdjointTerms::bytes(a.off.number_terms, true) + 15) & ~size_t(15);
  int const TD = a.diag.number_terms;
  double2 *d_v = reinterpret_cast<double2 *>(smem + p);
  double2 *chars = d_v + TD;
  uint64_t *d_m = reinterpret_cast<uint64_t *>(chars + a.number_chars);
  uint64_t *d_r = d_m + TD;
  uint64_t *d_s = d_r + TD;
  for (int t = threadIdx.x; t < TD; t += blockDim.x) {
    d_m[t] = a.diag.m[t];
    d_r[t] = a.diag.r[t];
    d_s[t] = a.diag.s[t];
    d_v[t] = a.diag.v[t];
  }
  for (int j = threadIdx.x; j < a.number_chars; j += blockDim.x) chars[j] = a.cvals[j];
  __syncthreads();

  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.chunk_rows) return;
  int64_t const row = a.chunk_begin + r;
  uint64_t const alpha = __ldg(a.rows + row);
  int const T = terms.T;
  bool const have_cidx = a.number_idx_planes > 0;
  double acc_r = 0.0, acc_i = 0.0;
  uint32_t q = __ldg(a.offsets + r);
  // Terms are matched 64 at a time into a bit mask (no divergence: every lane tests every
  // term), then each lane walks its own set bits -- the row's matrix elements in term order.
  for (int t0 = 0; t0 < T; t0 += 64) {
    int const tn = min(64, T - t0);
    uint64_t mask = 0;
    for (int k = 0; k < tn; ++k) mask |= (uint64_t)((alpha & terms.m[t0 + k]) == terms.l[t0 + k]) << k;
    while (mask != 0) {
      int const t = t0 + __ffsll((long long)mask) - 1;
      mask &= mask - 1;
      double2 w = terms.w[t];
      if (__popcll(alpha & terms.s[t]) & 1) { w.x = -w.x; w.y = -w.y; }
      double2 const ch = chars[have_cidx ? __ldcs(a.q_cidx + q) : 0];
      double const fr = ch.x * w.x + ch.y * w.y;  // conj(chi) * w
      double const fi = ch.x * w.y - ch.y * w.x;
      double2 xv = make_double2(0.0, 0.0);
      if (CPLX) xv = __ldcs(reinterpret_cast<double2 const *>(a.vals) + q);
      else xv.x = __ldcs(a.vals + q);
      ++q;
      if ((unsigned long long)__double_as_longlong(xv.x) == kMissBits) {
        // not in the basis although its norm is positive (DistributedMatrixVector.chpl:127-135)
        if (fr != 0.0 || fi != 0.0) atomicOr(a.error_flag, 1);
        continue;
      }
      if (CPLX) {
        acc_r += fr * xv.x - fi * xv.y;
        acc_i += fr * xv.y + fi * xv.x;
      } else {
        acc_r += fr * xv.x;
      }
    }
  }
  double dr = 0.0, di = 0.0;
  for (int k = 0; k < TD; ++k)
    if ((alpha & d_m[k]) == d_r[k]) {
      double const sign = (__popcll(alpha & d_s[k]) & 1) ? -1.0 : 1.0;
      dr += sign * d_v[k].x;
      di += sign * d_v[k].y;
    }
  double const ni = a.norms != nullptr ? __ldg(a.norms + row) : 1.0;
  int64_t const out = row - a.row_begin;
  if (CPLX) {
    double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.x) + row);
    double2 res;
    res.x = acc_r / ni + (dr * xv.x - di * xv.y);
    res.y = acc_i / ni + (dr * xv.y + di * xv.x);
    reinterpret_cast<double2 *>(a.y)[out] = res;
  } else {
    a.y[out] = acc_r / ni + dr * __ldg(a.x + row);  // kernels/reference.c:84-91 uses creal(v) only
  }
}

// Scalar fallback of orbit_kernel: thread per row, Benes walk per element.
__global__ void __launch_bounds__(128)
orbit_scalar_kernel(MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  AdjointTerms terms;
  terms.stage(smem, a.off, false);
  __syncthreads();
  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.chunk_rows) return;
  uint64_t const alpha = __ldg(a.rows + a.chunk_begin + r);
  uint32_t q = a.offsets[r];
  for (int t = 0; t < terms.T; ++t)
    if ((alpha & terms.m[t]) == terms.l[t]) {
      uint64_t rep;
      int e, flipped;
      orbit_min_global(a.g, alpha ^ terms.x[t], rep, e, flipped);
      unsigned c = 0;
      if (e >= 0) {
        unsigned const ci = __ldg(a.g.cinfo + e);
        c = flipped ? (ci >> 8) : (ci & 0xffu);
      }
      a.q_rep[q] = rep;
      a.q_cidx[q] = (uint8_t)c;
      ++q;
    }
}

// Thread per row: rank, gather, sum in term order, write y.
//   QUEUED: representatives / character indices come from the orbit kernel;
//   otherwise they are computed inline (no symmetries, or spin inversion only).
template <bool QUEUED, bool CPLX>
__global__ void __launch_bounds__(kGatherThreads, 8)  // <= 64 registers (the noinline norm check would take 128)
gather_kernel(MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  AdjointTerms terms;
  terms.stage(smem, a.off, true);
  size_t p = (AdjointTerms::bytes(a.off.number_terms, true) + 15) & ~size_t(15);
  int const TD = a.diag.number_terms;
  // 16-byte items first so that every array stays naturally aligned for any TD
  double2 *d_v = reinterpret_cast<double2 *>(smem + p);
  double2 *chars = d_v + TD;
  uint64_t *d_m = reinterpret_cast<uint64_t *>(chars + a.number_chars);
  uint64_t *d_r = d_m + TD;
  uint64_t *d_s = d_r + TD;
  for (int t = threadIdx.x; t < TD; t += blockDim.x) {
    d_m[t] = a.diag.m[t];
    d_r[t] = a.diag.r[t];
    d_s[t] = a.diag.s[t];
    d_v[t] = a.diag.v[t];
  }
  for (int j = threadIdx.x; j < a.number_chars; j += blockDim.x) chars[j] = a.cvals[j];
  __syncthreads();

  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.chunk_rows) return;
  int64_t const row = a.chunk_begin + r;
  uint64_t const alpha = __ldg(a.rows + row);
  IndexView const ix = a.ix;
  int const T = terms.T;
  double acc_r = 0.0, acc_i = 0.0;
  uint32_t q = QUEUED ? __ldg(a.offsets + r) : 0u;
  bool const skip_gather = (a.debug_skip & 2) != 0;

  int t = 0;
  while (t < T && !skip_gather) {
    // collect up to kGatherBatch matching terms
    uint64_t needle[kGatherBatch];
    double fr[kGatherBatch], fi[kGatherBatch];
    bool live[kGatherBatch];
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      while (t < T && (alpha & terms.m[t]) != terms.l[t]) ++t;
      live[u] = t < T;
      needle[u] = 0;
      fr[u] = fi[u] = 0.0;
      if (live[u]) {
        double2 w = terms.w[t];
        if (__popcll(alpha & terms.s[t]) & 1) { w.x = -w.x; w.y = -w.y; }
        uint64_t rep;
        unsigned c;
        if (QUEUED) {
          rep = __ldg(a.q_rep + q);
          c = a.number_idx_planes > 0 ? __ldg(a.q_cidx + q) : 0u;
          ++q;
        } else {
          rep = alpha ^ terms.x[t];
          c = 0;
          if (a.mode == kModeInversion) {
            // BatchedOperator.chpl:187-199
            uint64_t const inverted = rep ^ a.inversion_mask;
            if (inverted < rep) {
              rep = inverted;
              c = 1;
            }
          }
        }
        double2 const ch = chars[c];
        // conj(chi) * w
        fr[u] = ch.x * w.x + ch.y * w.y;
        fi[u] = ch.x * w.y - ch.y * w.x;
        needle[u] = rep;
        ++t;
      }
    }
    // rank: kGatherBatch independent branchless searches in lockstep
    int64_t j[kGatherBatch];
    if (ix.offsets32 != nullptr && !ix.identity) {  // lean 32-bit search on the basis' key type (warp-uniform choice)
      if (ix.lows16 != nullptr) index_find32<uint16_t, kGatherBatch>(ix, needle, live, j);
      else if (ix.lows32 != nullptr) index_find32<uint32_t, kGatherBatch>(ix, needle, live, j);
      else index_find32<uint64_t, kGatherBatch>(ix, needle, live, j);
    } else {
      index_find<kGatherBatch>(ix, needle, live, j);
    }
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      if (!live[u]) continue;
      if (j[u] >= 0) {
        if (CPLX) {
          double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.xs) + j[u]);
          acc_r += fr[u] * xv.x - fi[u] * xv.y;
          acc_i += fr[u] * xv.y + fi[u] * xv.x;
        } else {
          acc_r += fr[u] * __ldg(a.xs + j[u]);
        }
      } else if (fr[u] != 0.0 || fi[u] != 0.0) {
        // Not in the basis: fine when its norm vanishes (the reference
        // multiplies by n_beta = 0), an error otherwise
        // (DistributedMatrixVector.chpl:127-135).
        bool bad = true;
        if (a.mode == kModeGroup || a.mode == kModeGroupScalar)
          bad = stabiliser_sum_global(a.g, needle[u]) > kNormThreshold;
        if (bad) atomicOr(a.error_flag, 1);
      }
    }
  }

  // diagonal + write
  double dr = 0.0, di = 0.0;
  for (int k = 0; k < TD; ++k)
    if ((alpha & d_m[k]) == d_r[k]) {
      double const sign = (__popcll(alpha & d_s[k]) & 1) ? -1.0 : 1.0;
      dr += sign * d_v[k].x;
      di += sign * d_v[k].y;
    }
  double const ni = a.norms != nullptr ? __ldg(a.norms + row) : 1.0;
  int64_t const out = row - a.row_begin;
  if (CPLX) {
    double2 const xv = __ldg(reinterpret_cast<double2 const *>(a.x) + row);
    double2 res;
    res.x = acc_r / ni + (dr * xv.x - di * xv.y);
    res.y = acc_i / ni + (dr * xv.y + di * xv.x);
    reinterpret_cast<double2 *>(a.y)[out] = res;
  } else {
    // kernels/reference.c:84-91 uses creal(v) only
    a.y[out] = acc_r / ni + dr * __ldg(a.x + row);
  }
}

// xs[j] = n_j x[j]
__global__ void __launch_bounds__(256)
prescale_kernel(int64_t n, int complex_vectors, double const *__restrict__ norms,
                double const *__restrict__ x, double *__restrict__ xs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double const s = norms != nullptr ? norms[i] : 1.0;
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

// per_row[i] = matrix elements of row i (T < 32768, so 16 bits suffice)
__global__ void __launch_bounds__(256)
count_elements_rows_kernel(TermsView off, uint64_t const *__restrict__ reps, int64_t n, uint16_t *__restrict__ per_row) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const alpha = reps[i];
    unsigned c = 0;
    for (int t = 0; t < off.number_terms; ++t) c += ((alpha & __ldg(off.m + t)) == __ldg(off.l + t)) ? 1u : 0u;
    per_row[i] = (uint16_t)c;
  }
}
// out[b] = sum of per_row over [starts[b], starts[b + 1]) (one CTA per segment, round-robin: the work per row is a
// 2-byte read, so even a million-row segment takes microseconds)
__global__ void __launch_bounds__(256)
sum_segments_kernel(uint16_t const *__restrict__ per_row, uint64_t const *__restrict__ starts, int64_t number_segments,
                    uint64_t *__restrict__ out) {
  __shared__ unsigned long long warp_sums[8];
  for (int64_t b = blockIdx.x; b < number_segments; b += gridDim.x) {
    unsigned long long local = 0;
    for (uint64_t i = starts[b] + threadIdx.x; i < starts[b + 1]; i += blockDim.x) local += per_row[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long total = 0;
      for (int w = 0; w < 8; ++w) total += warp_sums[w];
      out[b] = total;
    }
    __syncthreads();
  }
}

static DeviceBuffer<uint16_t> &per_row_scratch() {
  static DeviceBuffer<uint16_t> b;
  return b;
}
void release_count_scratch(size_t above_bytes) {
  if (per_row_scratch().capacity * sizeof(uint16_t) > above_bytes) per_row_scratch().release();
}

// d_out[b] = matrix elements of rows [d_starts[b], d_starts[b + 1]); number_rows = d_starts[number_segments]
void count_elements_segments(OperatorDev &od, uint64_t const *d_rows, int64_t number_rows, uint64_t const *d_starts,
                             int64_t number_segments, uint64_t *d_out) {
  if (number_segments <= 0) return;
  Runtime &rt = runtime();
  uint16_t *d = per_row_scratch().reserve((size_t)number_rows + 1);
  if (number_rows > 0) {
    unsigned const blocks = (unsigned)std::min<int64_t>((number_rows + 255) / 256, (int64_t)rt.sm_count * 16);
    count_elements_rows_kernel<<<blocks, 256, 0, rt.stream>>>(od.off.view(), d_rows, number_rows, d);
  }
  unsigned const blocks = (unsigned)std::min<int64_t>(number_segments, (int64_t)rt.sm_count * 8);
  sum_segments_kernel<<<blocks, 256, 0, rt.stream>>>(d, d_starts, number_segments, d_out);
  count_launch(2);
  CUDA_CHECK(cudaGetLastError());
}

void ensure_norms(IndexData &ix, GroupData const &g);

void launch_prescale(int64_t n, bool complex_vectors, double const *norms, double const *x, double *xs,
                     cudaStream_t stream) {
  if (n <= 0) return;
  Runtime &rt = runtime();
  if (stream == nullptr) stream = rt.stream;
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 16);
  prescale_kernel<<<blocks, 256, 0, stream>>>(n, complex_vectors ? 1 : 0, norms, x, xs);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}


int64_t count_elements(OperatorDev &od, uint64_t const *d_rows, int64_t row_begin, int64_t row_end) {
  Runtime &rt = runtime();
  if (od.off.number_terms == 0 || row_end <= row_begin) return 0;
  static unsigned long long *d_counter = nullptr;
  if (d_counter == nullptr) CUDA_CHECK(cudaMalloc(&d_counter, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), rt.stream));
  int64_t const n = row_end - row_begin;
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
  count_elements_kernel<<<blocks, 256, 0, rt.stream>>>(od.off.view(), d_rows, row_begin, row_end, d_counter);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, d_counter, sizeof h, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  return (int64_t)h;
}

// Intermediates of one row chunk.  Two slots: with the pipelined split path chunk c + 1 is
// canonicalised (stream A) while chunk c is ranked and summed (stream B).
struct ChunkSlot {
  DeviceBuffer<uint32_t> counts, offsets;
  DeviceBuffer<uint64_t> q_rep;
  DeviceBuffer<uint8_t> q_cidx;
  DeviceBuffer<uint16_t> q_tsign;
  DeviceBuffer<double> vals;
  DeviceBuffer<uint64_t> q_sorted;  // sorted ranking
  DeviceBuffer<uint32_t> perm;
  cudaEvent_t orbit_done = nullptr, released = nullptr;
};
struct MatvecScratch {
  DeviceBuffer<double> x, xs, y;
  ChunkSlot slot[2];
  cudaStream_t stream_b = nullptr;
  cudaEvent_t inputs_ready = nullptr;
  // host-pointer entry point: finished row chunks of y drain to the caller's buffer on a copy stream while the
  // next chunks are computed
  // phased products: one slot per chunk, sized by the chunk's exact element count, kept between the phases
  std::vector<std::unique_ptr<ChunkSlot>> phase_slots;
  std::vector<int64_t> phase_elements;  // exact elements per chunk, cached with the tag below
  void const *phase_op = nullptr, *phase_index = nullptr;
  uint64_t phase_version = 0;  // OperatorDev::version the counts were taken with
  int64_t phase_row_begin = -1, phase_row_end = -1, phase_chunk_rows = -1;
  bool phase_ready = false;  // phase 1 has run for the tag and phase 2 has not consumed it yet
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_done;
  cudaEvent_t copies_done = nullptr, x_uploaded = nullptr;
  DeviceBuffer<unsigned char> scan_tmp;
  size_t scan_tmp_bytes = 0;
  DeviceBuffer<unsigned char> sort_tmp;  // sorted ranking: cub radix-sort scratch, positions 0, 1, 2, ...
  DeviceBuffer<uint32_t> iota;
  size_t iota_filled = 0;
  int *d_error = nullptr;
  double2 *d_plain_chars = nullptr;  // {1, +1, -1}: character table of the unprojected / inversion-only modes
  // per-kernel device time of the last matvec (LS_B200_PROFILE=1): events around every orbit / gather launch
  std::vector<cudaEvent_t> events;
  size_t events_used = 0;
  std::vector<std::pair<size_t, int>> spans;  // (index of start event, kind: 0 orbit, 1 gather, 2 row sum, 3 row count)
  // LS_B200_PROFILE, plain chunk loop: which chunk every span belongs to and the rows of every chunk -- the measured
  // cost density over the rows that the distributed driver re-balances the ranks with (ls_b200_dist_rebalance)
  std::vector<int> span_chunk;
  std::vector<std::pair<int64_t, int64_t>> chunk_rows;  // (first row relative to the call, rows)
  std::vector<double> chunk_ms;
  int64_t cost_rows = 0;  // rows of the call the costs belong to
};
static MatvecScratch &mv_scratch() {
  static MatvecScratch s;
  return s;
}

void matvec_forget_index(IndexData const *ix) {
  MatvecScratch &sc = mv_scratch();
  if (sc.phase_index == ix) {
    sc.phase_index = nullptr;
    sc.phase_op = nullptr;
    sc.phase_ready = false;
    sc.phase_elements.clear();
    sc.phase_slots.clear();  // the canonicalised elements of a basis that is going away: give the memory back
  }
}

// Frees the per-chunk buffers of the phased product (they are sized for one operator on one basis).
static void release_phase_slots() {
  MatvecScratch &sc = mv_scratch();
  sc.phase_slots.clear();
  sc.phase_elements.clear();
  sc.phase_op = nullptr;
  sc.phase_index = nullptr;
  sc.phase_ready = false;
}

__global__ void iota_kernel(uint32_t *__restrict__ out, size_t begin, size_t end) {
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (uint32_t)i;
}
__global__ void segment_starts_kernel(uint64_t *__restrict__ starts, int64_t number_segments, int64_t row_begin,
                                      int64_t row_end, int64_t segment_rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= number_segments; i += (int64_t)gridDim.x * blockDim.x)
    starts[i] = (uint64_t)min(row_end, row_begin + i * segment_rows);
}

// Chunks of rows holding at most `capacity` matrix elements each, cut at multiples of 16384 rows by exact counts
// (one counting pass, cached): the sorted-ranking path sizes its sorts and buffers by them.
static OperatorDev::ChunkPlan const &chunk_plan(OperatorDev &od, uint64_t const *d_rows, int64_t row_begin, int64_t row_end,
                                                int64_t capacity) {
  OperatorDev::ChunkPlan &p = od.plan;
  if (p.rows == d_rows && p.row_begin == row_begin && p.row_end == row_end && p.capacity == capacity && p.version == od.version)
    return p;
  Runtime &rt = runtime();
  int64_t const T = std::max(1, od.off.number_terms);
  int64_t segment_rows = 16384;
  while (segment_rows > 1 && segment_rows * T > capacity) segment_rows /= 2;  // a segment always fits
  int64_t const ns = (row_end - row_begin + segment_rows - 1) / segment_rows;
  static DeviceBuffer<uint64_t> starts, sums;
  uint64_t *d_starts = starts.reserve((size_t)ns + 1);
  uint64_t *d_sums = sums.reserve((size_t)ns + 1);
  segment_starts_kernel<<<(unsigned)std::min<int64_t>((ns + 256) / 256, 1024), 256, 0, rt.stream>>>(d_starts, ns, row_begin,
                                                                                                  row_end, segment_rows);
  count_launch();
  // (count_elements_segments indexes rows from 0: hand it the rows of this call)
  count_elements_segments(od, d_rows, row_end, d_starts, ns, d_sums);
  std::vector<uint64_t> h((size_t)ns);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), d_sums, sizeof(uint64_t) * (size_t)ns, cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  release_count_scratch(size_t(1) << 30);
  p.begin.assign(1, row_begin);
  p.elements.clear();
  int64_t acc = 0;
  for (int64_t i = 0; i < ns; ++i) {
    int64_t const n = (int64_t)h[(size_t)i];
    if (acc > 0 && acc + n > capacity) {
      p.begin.push_back(row_begin + i * segment_rows);
      p.elements.push_back(acc);
      acc = 0;
    }
    acc += n;
  }
  p.begin.push_back(row_end);
  p.elements.push_back(acc);
  p.rows = d_rows;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.capacity = capacity;
  p.version = od.version;
  return p;
}

template <class K>
static void allow_dynamic_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

static cudaEvent_t next_event(MatvecScratch &sc) {
  if (sc.events_used == sc.events.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    sc.events.push_back(e);
  }
  return sc.events[sc.events_used++];
}

// y[row_begin:row_end] = (H x)[row_begin:row_end]; x, y in device memory.
// number_vectors > 1 (block matvec, an extension: the reference halts, DistributedMatrixVector.chpl:1096-1097):
// vector v is x + v x_stride -> y + v y_stride (strides in scalars of the vector type).  On the split path all
// vectors share ONE canonicalisation + ranking pass -- the integer work is per matrix element, not per vector.
//
// target (distributed products, dist.cu): the rows are the caller's local shard (row numbers, x and y are LOCAL),
// while representatives are ranked against target->index -- the replicated index of the whole basis -- and the
// gathered values come from target->xs, the replicated pre-scaled vector n_j x_j.
void matvec_device(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, double const *d_x,
                   double *d_y, bool complex_vectors, int number_vectors, int64_t x_stride,
                   int64_t y_stride, double *host_y, int phase, MvTarget const *target) {
  // phase (split path, one vector): 1 = canonicalise only -- row counts, offsets, representatives of every chunk,
  // none of which depends on x -- into per-chunk buffers that persist; 2 = rank + gather + row sums from them.
  // A multi-GPU caller runs phase 1 of the NEXT product while NCCL all-gathers the result of this one.
  Runtime &rt = runtime();
  ls_hs_basis const *basis = op->basis;
  IndexData *ix = index_of(basis);
  LSB_CHECK(ix != nullptr, "basis is not built: call ls_hs_basis_build / ls_hs_build_representatives first");
  int64_t const dim = target != nullptr ? target->number_rows : ix->number_states;  // rows addressable by this call
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
  static bool const profile = getenv("LS_B200_PROFILE") != nullptr;
  sc.events_used = 0;
  sc.spans.clear();
  sc.span_chunk.clear();
  sc.chunk_rows.clear();
  sc.cost_rows = row_end - row_begin;
  CUDA_CHECK(cudaEventRecord(rt.ev0, rt.stream));

  MatvecArgs a{};
  a.ix = target != nullptr ? target->index : ix->view();
  a.rows = target != nullptr ? target->rows : ix->d_reps;
  a.off = od.off.view();
  a.diag = od.diag.view();
  a.complex_vectors = complex_vectors ? 1 : 0;
  a.row_begin = row_begin;
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
  if (info.has_permutation_symmetries) {
    GroupData const &g = *info.group;
    if (target == nullptr) ensure_norms(*ix, g);
    a.g = g.view();
    a.norms = target != nullptr ? target->norms : ix->d_norms;
    np = std::max(4, (g.number_bits + 3) / 4 * 4);
    inv = g.spin_inversion != 0;
    char const *env = getenv("LS_B200_MATVEC");
    bool const want_scalar = env != nullptr && strcmp(env, "scalar") == 0;
    LSB_CHECK(!g.cinfo.empty(), "symmetry sectors with more than 256 distinct character values are not supported");
    a.cvals = g.d_cvals;
    a.number_chars = (int)(g.cvals.size() / 2);
    while ((1 << a.number_idx_planes) < a.number_chars) ++a.number_idx_planes;
    bool const bitsliced = !want_scalar && orbit_prepare(g, np);
    a.mode = bitsliced ? kModeGroup : kModeGroupScalar;
    if (!bitsliced) a.number_idx_planes = std::max(a.number_idx_planes, 1);  // the scalar kernel always writes q_cidx
    if (target == nullptr) {
      // pre-scaled copy of x
      size_t const words = (size_t)dim * (complex_vectors ? 2 : 1);
      double *xs = sc.xs.reserve(words * (size_t)number_vectors);
      unsigned const blocks = (unsigned)std::min<int64_t>((dim + 255) / 256, (int64_t)rt.sm_count * 16);
      for (int v = 0; v < number_vectors && phase != 1; ++v) {
        // (xs is packed: vector v at xs + v dim, whatever the caller's stride)
        prescale_kernel<<<blocks, 256, 0, rt.stream>>>(dim, a.complex_vectors, ix->d_norms,
                                                       d_x + (size_t)v * (size_t)x_stride * (complex_vectors ? 2 : 1),
                                                       xs + (size_t)v * words);
        count_launch();
      }
      CUDA_CHECK(cudaGetLastError());
      a.xs = xs;
    }
  } else if (info.has_spin_inversion) {
    a.mode = kModeInversion;
    // cidx 1 = the spin-inversion character: entry 1 of the plain table is +1, entry 2 is -1
    a.cvals = sc.d_plain_chars + (basis->spin_inversion < 0 ? 1 : 0);
    a.number_chars = 2;
  }
  if (target != nullptr) {
    LSB_CHECK(phase == 0 && number_vectors == 1, "distributed products take one vector at a time");
    a.xs = target->xs;  // the replicated vector (already multiplied by the norms where the basis has them)
  }
  LSB_CHECK(a.off.number_terms < 0x8000, "too many off-diagonal terms");

  int const T = a.off.number_terms;
  bool const queued = (a.mode == kModeGroup || a.mode == kModeGroupScalar) && T > 0;
  size_t const gather_smem = ((AdjointTerms::bytes(T, true) + 15) & ~size_t(15)) + (size_t)a.diag.number_terms * 40 +
                             (size_t)a.number_chars * 16;
  size_t const count_smem = AdjointTerms::bytes(T, false);
  size_t orbit_smem = ((count_smem + 15) & ~size_t(15)) + (size_t)(kOrbitThreads / 32) * kWarpSlabBytes;
  size_t const fused_smem = ((AdjointTerms::bytes(T, true) + 15) & ~size_t(15)) + (size_t)a.number_chars * 16 +
                            (size_t)(kOrbitThreads / 32) * kFusedWarpBytes;
  size_t const sum_smem = (size_t)a.diag.number_terms * 40;
  // Pipeline variants (LS_B200_MATVEC, for A/B measurements; results agree to rounding):
  //   split   (default) orbit kernel -> full-occupancy rank + gather kernel -> per-row sum
  //   fused   canonicalise + rank + gather in one kernel, then a per-row sum
  //   unfused orbit kernel -> thread-per-row rank + gather + sum (the first design)
  //   scalar  split, with the scalar Benes walk instead of the bit-sliced orbit kernel (handled above)
  char const *variant = getenv("LS_B200_MATVEC");
  bool const is_group = a.mode == kModeGroup || a.mode == kModeGroupScalar;
  bool const want_fused = variant != nullptr && strcmp(variant, "fused") == 0;
  bool const want_unfused = variant != nullptr && strcmp(variant, "unfused") == 0;
  bool const fused = a.mode == kModeGroup && want_fused && fused_smem <= rt.smem_optin && sum_smem <= rt.smem_optin;
  bool const split = is_group && !fused && !want_unfused && sum_smem <= rt.smem_optin;
  if (fused) orbit_smem = fused_smem;
  // the orbit kernel can carry (term, sign) in the spare bits of the staged states only when NP <= 48
  bool const want_tsign = split && a.mode == kModeGroup && np <= 48;
  if (want_tsign)
    orbit_smem = ((AdjointTerms::bytes(T, true) + 15) & ~size_t(15)) + (size_t)(kOrbitThreads / 32) * kWarpSlabBytes;
  LSB_CHECK(gather_smem <= rt.smem_optin && orbit_smem <= rt.smem_optin,
            "operator / symmetry tables do not fit in shared memory");
  a.number_vectors = 1;
  if (number_vectors > 1) {
    if (!(split && want_tsign && T > 0)) {
      // no shared pass to amortise on this path: one vector at a time
      size_t const scalar = complex_vectors ? 2 : 1;
      for (int v = 0; v < number_vectors; ++v)
        matvec_device(op, row_begin, row_end, d_x + (size_t)v * (size_t)x_stride * scalar,
                      d_y + (size_t)v * (size_t)y_stride * scalar, complex_vectors, 1, 0, 0,
                      host_y != nullptr ? host_y + (size_t)v * (size_t)y_stride * scalar : nullptr);
      return;
    }
    a.number_vectors = number_vectors;
    a.x_stride = x_stride;
    a.xs_stride = dim;
    a.y_stride = y_stride;
  }

  // Chunk of rows: its intermediates (up to 19 bytes per matrix element) must fit the
  // scratch capacity even if every term matched every row.
  int64_t capacity = (int64_t(1) << 27) / a.number_vectors;
  if (char const *env = getenv("LS_B200_MV_CHUNK")) capacity = std::max<int64_t>(4096, atoll(env));
  // Pipelined split path (experimental): two half-size slots, rank + gather + row sum of
  // chunk c on a second stream while the orbit kernel of chunk c + 1 runs on the first -- the two are bound by
  // different resources (memory latency vs integer issue).
  // Measured on kagome-36: no gain (89 ms vs 86.7 ms) -- both kernels fill the machine on their own, the hardware
  // runs them back to back -- so it is off unless LS_B200_MV_PIPELINE=1.
  bool pipelined = false;
  if (char const *env = getenv("LS_B200_MV_PIPELINE")) pipelined = split && T > 0 && atoi(env) != 0;
  if (pipelined && (row_end - row_begin) * (int64_t)T <= capacity / 2) pipelined = false;  // a single chunk
  if (pipelined) capacity /= 2;
  int64_t chunk_rows = row_end - row_begin;
  bool const phased = phase != 0 && split && T > 0 && a.mode == kModeGroup && a.number_vectors == 1;
  if (phased) pipelined = false;
  int const number_slots = phased ? 0 : (pipelined ? 2 : 1);
  // Sorted ranking (LS_B200_MV_SORT=1, off by default).  The chunk's representatives are radix-sorted by their
  // leading bits before the index search, so that consecutive threads walk the level-1 table, the keys and the vector
  // in ascending order, and the values are scattered back to CSR order; chunks are cut by exact element counts.
  // Bit-identical results.  Measured on a 17 GB footprint (chain-40, translations only: 3.5e10 elements): the sort
  // costs 2.0 s against a 0.76 s unsorted rank + gather, whose random accesses the memory system sustains as long as
  // the tables are mapped with large pages (see alloc_local) -- so it stays an A/B knob.
  bool sorted = false;
  if (split && queued && want_tsign && a.number_vectors == 1 && phase == 0 && !pipelined && a.mode == kModeGroup) {
    if (char const *env = getenv("LS_B200_MV_SORT")) sorted = atoi(env) != 0;
  }
  OperatorDev::ChunkPlan const *plan = nullptr;
  int sort_begin_bit = 0, sort_end_bit = 64;
  if (sorted) {
    plan = &chunk_plan(od, a.rows, row_begin, row_end, capacity);
    int sort_bits = 24;
    if (char const *env = getenv("LS_B200_MV_SORT_BITS")) sort_bits = std::max(1, atoi(env));
    sort_end_bit = std::max(1, info.number_bits);
    sort_begin_bit = std::max(0, sort_end_bit - sort_bits);
  }
  if (queued) {
    if (sorted) {
      chunk_rows = 1;
      for (size_t c = 0; c + 1 < plan->begin.size(); ++c) chunk_rows = std::max(chunk_rows, plan->begin[c + 1] - plan->begin[c]);
    } else {
      chunk_rows = std::max<int64_t>(1, std::min<int64_t>(chunk_rows, capacity / T));
      capacity = chunk_rows * T;
    }
    for (int k = 0; k < number_slots; ++k) {
      ChunkSlot &slot = sc.slot[k];
      slot.counts.reserve((size_t)chunk_rows + 1);
      slot.offsets.reserve((size_t)chunk_rows + 1);
      if (sorted) {
        slot.q_sorted.reserve((size_t)capacity + 32);
        slot.perm.reserve((size_t)capacity + 32);
      }
      if (split || fused)
        slot.vals.reserve(((size_t)capacity + 32) * (complex_vectors ? 2 : 1) * (size_t)a.number_vectors);
      if (!fused) {
        slot.q_rep.reserve((size_t)capacity + 32);
        slot.q_cidx.reserve((size_t)capacity + 32);
        if (want_tsign) slot.q_tsign.reserve((size_t)capacity + 32);
      }
      if (pipelined && slot.orbit_done == nullptr) {
        CUDA_CHECK(cudaEventCreateWithFlags(&slot.orbit_done, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&slot.released, cudaEventDisableTiming));
      }
    }
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(chunk_rows + 1), rt.stream);
    if (tmp > sc.scan_tmp_bytes) {
      sc.scan_tmp.reserve(tmp);
      sc.scan_tmp_bytes = sc.scan_tmp.capacity;
    }
    if (sorted) {
      size_t bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t const *)nullptr, (uint64_t *)nullptr, (uint32_t const *)nullptr,
                                      (uint32_t *)nullptr, (int)capacity, sort_begin_bit, sort_end_bit, rt.stream);
      sc.sort_tmp.reserve(bytes);
      uint32_t *iota = sc.iota.reserve((size_t)capacity + 32);
      if (sc.iota_filled < (size_t)capacity || sc.iota.capacity != sc.iota_filled) {
        iota_kernel<<<rt.sm_count * 8, 256, 0, rt.stream>>>(iota, 0, sc.iota.capacity);
        count_launch();
        sc.iota_filled = sc.iota.capacity;
      }
    }
    if (a.mode == kModeGroup) {
      LSB_CHECK(np >= 4 && np <= 64 && (np & 3) == 0, "unsupported number of bits");
    } else {
      allow_dynamic_smem(orbit_scalar_kernel, count_smem);
    }
    allow_dynamic_smem(row_count_kernel, count_smem);
  } else {
    chunk_rows = std::min<int64_t>(chunk_rows, int64_t(1) << 30);
  }
  auto gather = queued ? (complex_vectors ? gather_kernel<true, true> : gather_kernel<true, false>)
                       : (complex_vectors ? gather_kernel<false, true> : gather_kernel<false, false>);
  allow_dynamic_smem(gather, gather_smem);
  auto row_sum = complex_vectors ? row_sum_kernel<true> : row_sum_kernel<false>;
  if (fused || split) allow_dynamic_smem(row_sum, sum_smem);
  auto row_combine = complex_vectors ? row_combine_kernel<true> : row_combine_kernel<false>;
  if (split) allow_dynamic_smem(row_combine, gather_smem);
  void (*rank_gather)(MatvecArgs) = rank_gather_kernel<void>;
  if (a.ix.offsets32 != nullptr && !a.ix.identity)
    rank_gather = a.ix.lows16 != nullptr   ? rank_gather_kernel<uint16_t>
                  : a.ix.lows32 != nullptr ? rank_gather_kernel<uint32_t>
                                           : rank_gather_kernel<uint64_t>;
  else if (a.ix.offsets64 != nullptr && !a.ix.identity && (a.ix.lows16 != nullptr || a.ix.lows32 != nullptr))
    // 2^32 states or more (the replicated index of a distributed basis): 64-bit bucket starts, 32-bit windows
    rank_gather = a.ix.lows16 != nullptr ? rank_gather_kernel<uint16_t, true> : rank_gather_kernel<uint32_t, true>;
  // the same kernels reading the chunk's representatives in sorted order (LS_B200_MV_SORT=1)
  void (*rank_gather_sorted)(MatvecArgs) = rank_gather_kernel<void, false, true>;
  if (a.ix.offsets32 != nullptr && !a.ix.identity)
    rank_gather_sorted = a.ix.lows16 != nullptr   ? rank_gather_kernel<uint16_t, false, true>
                         : a.ix.lows32 != nullptr ? rank_gather_kernel<uint32_t, false, true>
                                                  : rank_gather_kernel<uint64_t, false, true>;
  else if (a.ix.offsets64 != nullptr && !a.ix.identity && (a.ix.lows16 != nullptr || a.ix.lows32 != nullptr))
    rank_gather_sorted = a.ix.lows16 != nullptr ? rank_gather_kernel<uint16_t, true, true> : rank_gather_kernel<uint32_t, true, true>;

  // rank_gather runs as a persistent grid sized to the machine (tables staged once, no empty CTAs: 8 % faster);
  // the orbit kernel keeps one CTA per four blocks (see there).
  unsigned rank_resident = ~0u;
  if (split && queued) {
    int per_sm = 0;
    size_t const rank_smem = want_tsign ? ((size_t)T + (size_t)a.number_chars) * 16 : 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rank_gather, kRankThreads, rank_smem));
    rank_resident = (unsigned)std::max(1, per_sm) * (unsigned)rt.sm_count;
  }
  // host-pointer callers: rows [begin, begin + nrows) of y are final once `producer` reaches this point -- copy them
  // out on the copy stream behind the next chunk's kernels
  auto drain = [&](int64_t chunk_no, int64_t begin, int64_t nrows, cudaStream_t producer) {
    if (host_y == nullptr) return;
    if (sc.copy_stream == nullptr) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&sc.copy_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&sc.copies_done, cudaEventDisableTiming));
    }
    while (sc.chunk_done.size() <= (size_t)chunk_no) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      sc.chunk_done.push_back(e);
    }
    CUDA_CHECK(cudaEventRecord(sc.chunk_done[(size_t)chunk_no], producer));
    CUDA_CHECK(cudaStreamWaitEvent(sc.copy_stream, sc.chunk_done[(size_t)chunk_no], 0));
    size_t const scalar = sizeof(double) * (complex_vectors ? 2 : 1);
    for (int v = 0; v < a.number_vectors; ++v) {
      size_t const at = ((size_t)v * (size_t)(a.number_vectors > 1 ? a.y_stride : 0) + (size_t)(begin - row_begin)) * scalar;
      CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(host_y) + at, reinterpret_cast<char const *>(d_y) + at,
                                 (size_t)nrows * scalar, cudaMemcpyDeviceToHost, sc.copy_stream));
    }
  };
  auto drain_finish = [&]() {
    if (host_y == nullptr || sc.copy_stream == nullptr) return;
    CUDA_CHECK(cudaEventRecord(sc.copies_done, sc.copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(rt.stream, sc.copies_done, 0));
  };
  if (phase == 1 && !phased) return;  // nothing to precompute on this path: phase 2 does the whole product
  if (phased) {
    int64_t const number_chunks = (row_end - row_begin + chunk_rows - 1) / chunk_rows;
    bool const same = sc.phase_op == op && sc.phase_index == ix && sc.phase_version == od.version &&
                      sc.phase_row_begin == row_begin &&
                      sc.phase_row_end == row_end && sc.phase_chunk_rows == chunk_rows &&
                      (int64_t)sc.phase_elements.size() == number_chunks;
    if (!same) {
      // exact element counts per chunk (one counting pass, cached): the persistent buffers are sized by them
      sc.phase_elements.assign((size_t)number_chunks, 0);
      for (int64_t c = 0; c < number_chunks; ++c) {
        int64_t const begin = row_begin + c * chunk_rows;
        sc.phase_elements[(size_t)c] = count_elements(od, a.rows, begin, std::min(row_end, begin + chunk_rows));
      }
      sc.phase_op = op;
      sc.phase_index = ix;
      sc.phase_version = od.version;
      sc.phase_row_begin = row_begin;
      sc.phase_row_end = row_end;
      sc.phase_chunk_rows = chunk_rows;
      sc.phase_ready = false;
      while ((int64_t)sc.phase_slots.size() < number_chunks) sc.phase_slots.push_back(std::make_unique<ChunkSlot>());
    }
    int64_t most = 0;
    for (int64_t n : sc.phase_elements) most = std::max(most, n);
    auto bind = [&](int64_t c) {
      ChunkSlot &slot = *sc.phase_slots[(size_t)c];
      int64_t const begin = row_begin + c * chunk_rows;
      int64_t const nrows = std::min(chunk_rows, row_end - begin);
      size_t const n = (size_t)sc.phase_elements[(size_t)c] + 32;
      a.chunk_begin = begin;
      a.chunk_rows = (int)nrows;
      a.counts = slot.counts.reserve((size_t)nrows + 1);
      a.offsets = slot.offsets.reserve((size_t)nrows + 1);
      a.q_rep = slot.q_rep.reserve(n);
      a.q_cidx = slot.q_cidx.reserve(n);
      a.q_tsign = want_tsign ? slot.q_tsign.reserve(n) : nullptr;
      a.vals = sc.slot[0].vals.reserve(((size_t)most + 32) * (complex_vectors ? 2 : 1));
      a.vals_stride = most + 32;
      return nrows;
    };
    if (phase == 1 || !sc.phase_ready) {
      for (int64_t c = 0; c < number_chunks; ++c) {
        int64_t const nrows = bind(c);
        row_count_kernel<<<ceil_div((size_t)nrows + 1, 256), 256, count_smem, rt.stream>>>(a);
        size_t tmp = sc.scan_tmp_bytes;
        cub::DeviceScan::ExclusiveSum(sc.scan_tmp.ptr, tmp, a.counts, a.offsets, (int)(nrows + 1), rt.stream);
        size_t const words = (((size_t)sc.phase_elements[(size_t)c] + 1023) / 1024) * 32;
        if (profile) {
          sc.spans.emplace_back(sc.events_used, 0);
          CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
        }
        orbit_launch(np, inv, fused, words, orbit_smem, rt.stream, a);
        if (profile) CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
        count_launch(3);
      }
      CUDA_CHECK(cudaGetLastError());
      sc.phase_ready = true;
    }
    if (phase == 2) {
      for (int64_t c = 0; c < number_chunks; ++c) {
        int64_t const nrows = bind(c);
        size_t const tiles = ceil_div((size_t)sc.phase_elements[(size_t)c], 32 * kRankBatch);
        size_t const rank_smem = a.q_tsign != nullptr ? ((size_t)T + (size_t)a.number_chars) * 16 : 0;
        if (profile) {
          sc.spans.emplace_back(sc.events_used, 1);
          CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
        }
        if (tiles > 0) {
          unsigned const blocks = std::min<unsigned>(ceil_div(tiles, kRankThreads / 32), rank_resident);
          rank_gather<<<blocks, kRankThreads, rank_smem, rt.stream>>>(a);
        }
        if (profile) {
          CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
          sc.spans.emplace_back(sc.events_used, 2);
          CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
        }
        if (a.q_tsign != nullptr)
          row_sum<<<ceil_div((size_t)nrows, 256), 256, sum_smem, rt.stream>>>(a);
        else
          row_combine<<<ceil_div((size_t)nrows, kGatherThreads), kGatherThreads, gather_smem, rt.stream>>>(a);
        if (profile) CUDA_CHECK(cudaEventRecord(next_event(sc), rt.stream));
        count_launch(2);
        drain(c, a.chunk_begin, nrows, rt.stream);
      }
      CUDA_CHECK(cudaGetLastError());
      drain_finish();
      sc.phase_ready = false;
    }
    CUDA_CHECK(cudaEventRecord(rt.ev1, rt.stream));
    return;
  }
  cudaStream_t const stream_a = rt.stream;
  cudaStream_t stream_b = rt.stream;
  if (pipelined) {
    if (sc.stream_b == nullptr) {
      int least = 0, greatest = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      int prio = greatest;  // rank + gather first: its CTAs are short and few of them fit next to the orbit kernel's
      if (char const *env = getenv("LS_B200_MV_PRIORITY")) prio = atoi(env) != 0 ? greatest : least;
      CUDA_CHECK(cudaStreamCreateWithPriority(&sc.stream_b, cudaStreamNonBlocking, prio));
      CUDA_CHECK(cudaEventCreateWithFlags(&sc.inputs_ready, cudaEventDisableTiming));
    }
    stream_b = sc.stream_b;
    // stream B reads x, xs (and everything earlier work on the library stream produced)
    CUDA_CHECK(cudaEventRecord(sc.inputs_ready, stream_a));
    CUDA_CHECK(cudaStreamWaitEvent(stream_b, sc.inputs_ready, 0));
  }

  int64_t chunk_index = 0;
  for (int64_t begin = row_begin; begin < row_end; ++chunk_index) {
    int64_t const nrows = sorted ? plan->begin[(size_t)chunk_index + 1] - begin : std::min(chunk_rows, row_end - begin);
    ChunkSlot &slot = sc.slot[pipelined ? (chunk_index & 1) : 0];
    if (profile) sc.chunk_rows.emplace_back(begin - row_begin, nrows);
    a.chunk_begin = begin;
    a.chunk_rows = (int)nrows;
    a.counts = slot.counts.ptr;
    a.offsets = slot.offsets.ptr;
    a.q_rep = slot.q_rep.ptr;
    a.q_cidx = slot.q_cidx.ptr;
    a.q_tsign = want_tsign ? slot.q_tsign.ptr : nullptr;
    a.vals = slot.vals.ptr;
    a.vals_stride = capacity + 32;
    if (queued) {
      if (pipelined && chunk_index >= 2) CUDA_CHECK(cudaStreamWaitEvent(stream_a, slot.released, 0));
      if (profile) {
        sc.spans.emplace_back(sc.events_used, 3);
        sc.span_chunk.push_back((int)sc.chunk_rows.size() - 1);
        CUDA_CHECK(cudaEventRecord(next_event(sc), stream_a));
      }
      row_count_kernel<<<ceil_div((size_t)nrows + 1, 256), 256, count_smem, stream_a>>>(a);
      size_t tmp = sc.scan_tmp_bytes;
      cub::DeviceScan::ExclusiveSum(sc.scan_tmp.ptr, tmp, a.counts, a.offsets, (int)(nrows + 1), stream_a);
      count_launch(2);
      if (profile) {
        CUDA_CHECK(cudaEventRecord(next_event(sc), stream_a));
        sc.spans.emplace_back(sc.events_used, 0);
        sc.span_chunk.push_back((int)sc.chunk_rows.size() - 1);
        CUDA_CHECK(cudaEventRecord(next_event(sc), stream_a));
      }
      if (a.mode == kModeGroup) {
        // grid for the worst case (every term matches); words past the chunk's total exit at once
        size_t const max_words = (((size_t)nrows * (size_t)T + 1023) / 1024) * 32;
        orbit_launch(np, inv, fused, max_words, orbit_smem, stream_a, a);
      } else {
        orbit_scalar_kernel<<<ceil_div((size_t)nrows, 128), 128, count_smem, stream_a>>>(a);
      }
      count_launch();
      if (profile) CUDA_CHECK(cudaEventRecord(next_event(sc), stream_a));
      if (pipelined) {
        CUDA_CHECK(cudaEventRecord(slot.orbit_done, stream_a));
        CUDA_CHECK(cudaStreamWaitEvent(stream_b, slot.orbit_done, 0));
      }
    }
    if (chunk_index == 0 && target != nullptr && target->xs_ready != nullptr)
      CUDA_CHECK(cudaStreamWaitEvent(stream_b, target->xs_ready, 0));  // the replicated vector is read from here on
    if (profile) {
      sc.spans.emplace_back(sc.events_used, 1);
      sc.span_chunk.push_back((int)sc.chunk_rows.size() - 1);
      CUDA_CHECK(cudaEventRecord(next_event(sc), stream_b));
    }
    if (fused && queued) {
      row_sum<<<ceil_div((size_t)nrows, 256), 256, sum_smem, stream_b>>>(a);
    } else if (split && queued) {
      size_t max_tiles = ceil_div((size_t)nrows * (size_t)T, 32 * kRankBatch);
      a.perm = nullptr;
      a.q_sorted = nullptr;
      if (sorted) {
        int64_t const n = plan->elements[(size_t)chunk_index];
        LSB_CHECK(n <= capacity, "chunk plan exceeds the buffer capacity");
        max_tiles = ceil_div((size_t)std::max<int64_t>(n, 1), 32 * kRankBatch);
        if (n > 0) {
          size_t bytes = sc.sort_tmp.capacity;
          cub::DeviceRadixSort::SortPairs(sc.sort_tmp.ptr, bytes, (uint64_t const *)a.q_rep, slot.q_sorted.ptr,
                                          (uint32_t const *)sc.iota.ptr, slot.perm.ptr, (int)n, sort_begin_bit, sort_end_bit,
                                          stream_b);
          count_launch(4);
          a.perm = slot.perm.ptr;
          a.q_sorted = slot.q_sorted.ptr;
        }
      }
      size_t const rank_smem = a.q_tsign != nullptr ? ((size_t)T + (size_t)a.number_chars) * 16 : 0;
      unsigned const blocks = std::min<unsigned>(ceil_div(max_tiles, kRankThreads / 32), rank_resident);
      if (a.perm != nullptr) rank_gather_sorted<<<blocks, kRankThreads, rank_smem, stream_b>>>(a);
      else rank_gather<<<blocks, kRankThreads, rank_smem, stream_b>>>(a);
      count_launch();
      if (profile) {
        CUDA_CHECK(cudaEventRecord(next_event(sc), stream_b));
        sc.spans.emplace_back(sc.events_used, 2);
        sc.span_chunk.push_back((int)sc.chunk_rows.size() - 1);
        CUDA_CHECK(cudaEventRecord(next_event(sc), stream_b));
      }
      if (a.q_tsign != nullptr)
        row_sum<<<dim3(ceil_div((size_t)nrows, 256), (unsigned)a.number_vectors), 256, sum_smem, stream_b>>>(a);
      else
        row_combine<<<ceil_div((size_t)nrows, kGatherThreads), kGatherThreads, gather_smem, stream_b>>>(a);
    } else
      gather<<<ceil_div((size_t)nrows, kGatherThreads), kGatherThreads, gather_smem, stream_b>>>(a);
    count_launch();
    if (profile) CUDA_CHECK(cudaEventRecord(next_event(sc), stream_b));
    if (pipelined) CUDA_CHECK(cudaEventRecord(slot.released, stream_b));
    CUDA_CHECK(cudaGetLastError());
    drain(chunk_index, begin, nrows, stream_b);
    begin += nrows;
  }
  drain_finish();
  CUDA_CHECK(cudaEventRecord(rt.ev1, rt.stream));
}

// Blocks until the stream drains, then reports kernel time and sector errors.
bool matvec_finish() {
  Runtime &rt = runtime();
  MatvecScratch &sc = mv_scratch();
  int flag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&flag, sc.d_error, sizeof(int), cudaMemcpyDeviceToHost, rt.stream));
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, rt.ev0, rt.ev1) == cudaSuccess) rt.last_matvec_ms = ms;
  rt.last_orbit_ms = rt.last_gather_ms = rt.last_combine_ms = rt.last_count_ms = 0;
  rt.last_orbit_launches = rt.last_gather_launches = 0;
  bool const by_chunk = !sc.spans.empty() && sc.span_chunk.size() == sc.spans.size();
  sc.chunk_ms.assign(by_chunk ? sc.chunk_rows.size() : 0, 0.0);
  size_t span_no = 0;
  for (auto const &span : sc.spans) {
    float t = 0;
    size_t const this_span = span_no++;
    if (cudaEventElapsedTime(&t, sc.events[span.first], sc.events[span.first + 1]) != cudaSuccess) continue;
    if (by_chunk && sc.span_chunk[this_span] >= 0) sc.chunk_ms[(size_t)sc.span_chunk[this_span]] += t;
    if (span.second == 0) { rt.last_orbit_ms += t; ++rt.last_orbit_launches; }
    else if (span.second == 2) { rt.last_combine_ms += t; }
    else if (span.second == 3) { rt.last_count_ms += t; }
    else { rt.last_gather_ms += t; ++rt.last_gather_launches; }
  }
  if (flag != 0) {
    CUDA_CHECK(cudaMemsetAsync(sc.d_error, 0, sizeof(int), rt.stream));
    return false;
  }
  return true;
}

// Measured cost (ms of kernel time, LS_B200_PROFILE) of the rows of the last product, resampled to `segments` equal
// pieces of its row range; false when the last product recorded none.
bool matvec_last_row_costs(int64_t number_rows, int segments, double *out) {
  MatvecScratch &sc = mv_scratch();
  for (int k = 0; k < segments; ++k) out[k] = 0.0;
  if (sc.chunk_ms.empty() || sc.chunk_ms.size() != sc.chunk_rows.size() || sc.cost_rows != number_rows || number_rows <= 0)
    return false;
  for (size_t c = 0; c < sc.chunk_ms.size(); ++c) {
    int64_t const b = sc.chunk_rows[c].first, n = sc.chunk_rows[c].second;
    if (n <= 0) continue;
    // spread the chunk's time over the segments it overlaps, in proportion to the rows
    int64_t const k0 = (int64_t)(((__int128)b * segments) / number_rows);
    int64_t const k1 = (int64_t)(((__int128)(b + n - 1) * segments) / number_rows);
    for (int64_t k = k0; k <= k1 && k < segments; ++k) {
      int64_t const lo = std::max<int64_t>(b, (int64_t)(((__int128)k * number_rows + segments - 1) / segments));
      int64_t const hi = std::min<int64_t>(b + n, (int64_t)(((__int128)(k + 1) * number_rows + segments - 1) / segments));
      if (hi > lo) out[k] += sc.chunk_ms[c] * (double)(hi - lo) / (double)n;
    }
  }
  return true;
}

char const *kInvalidIndexMessage =
    "matrix_vector_product: the operator maps a basis state outside of the basis with a non-zero "
    "coefficient (invalid index); it does not respect the symmetries of the basis";


// ---- push form: all-to-all products of a basis sharded over ranks -----------------------------------------
// The rank that owns COLUMN i applies the forward terms to alpha_i, canonicalises every beta (same orbit kernel as the
// pull form) and emits records (rep_j, chi_g v_t sign x_i / n_i) grouped by the rank that owns rep_j
// (BatchedOperator.chpl:207-253 for the coefficient; owner = contiguous range instead of the hash of
// StatesEnumeration.chpl:198-212); the owner ranks the representative locally and adds n_j times the coefficient
// into y with fp64 atomics (ConcurrentAccessor.chpl:31-33).
constexpr int kPushThreads = 256;
constexpr int kPushMaxWorld = 64;

struct PushArgs {
  MatvecArgs a;             // rows, chunk, offsets, q_rep, q_cidx; a.off = the operator's terms (forward use: match on r)
  double const *x;          // local vector (row numbering of a.rows)
  uint64_t const *splitters;  // [world] first representative of every rank
  int world;
  int nblocks;
  uint32_t *hist;           // [world][nblocks] records per (owner, block)
  uint32_t const *base;     // exclusive scan of hist
  PushRecord *send;
};

__device__ __forceinline__ int push_owner(uint64_t const *splitters, int world, uint64_t rep) {
  int d = 0;
  for (int k = 1; k < world; ++k) d += (splitters[k] <= rep) ? 1 : 0;  // splitters ascend; empty ranks carry ~0
  return d;
}

// Thread per column.  SCATTER = false: count the records per owner; true: write them.
template <bool SCATTER>
__global__ void __launch_bounds__(kPushThreads)
push_emit_kernel(PushArgs const p) {
  extern __shared__ __align__(16) unsigned char smem[];
  MatvecArgs const &a = p.a;
  int const T = a.off.number_terms;
  double2 *t_v = reinterpret_cast<double2 *>(smem);
  uint64_t *t_m = reinterpret_cast<uint64_t *>(t_v + T);
  uint64_t *t_r = t_m + T;
  uint64_t *t_x = t_r + T;
  uint64_t *t_s = t_x + T;
  double2 *chars = reinterpret_cast<double2 *>(t_s + T);
  uint64_t *split = reinterpret_cast<uint64_t *>(chars + a.number_chars);
  unsigned *counters = reinterpret_cast<unsigned *>(split + p.world);
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    t_v[t] = a.off.v[t];
    t_m[t] = a.off.m[t];
    t_r[t] = a.off.r[t];
    t_x[t] = a.off.x[t];
    t_s[t] = a.off.s[t];
  }
  for (int j = threadIdx.x; j < a.number_chars; j += blockDim.x) chars[j] = a.cvals[j];
  for (int d = threadIdx.x; d < p.world; d += blockDim.x) {
    split[d] = p.splitters[d];
    counters[d] = SCATTER ? p.base[(size_t)d * p.nblocks + blockIdx.x] : 0u;
  }
  __syncthreads();
  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < a.chunk_rows) {
    int64_t const row = a.chunk_begin + r;
    uint64_t const alpha = __ldg(a.rows + row);
    double const xi = __ldg(p.x + row);
    double const scale = a.norms != nullptr ? xi / __ldg(a.norms + row) : xi;
    bool const queued = a.mode == kModeGroup || a.mode == kModeGroupScalar;
    uint32_t q = queued ? __ldg(a.offsets + r) : 0u;
    if (scale != 0.0) {
      for (int t = 0; t < T; ++t) {
        if ((alpha & t_m[t]) != t_r[t]) continue;
        uint64_t rep;
        unsigned c = 0;
        if (queued) {
          rep = __ldg(a.q_rep + q);
          c = a.number_idx_planes > 0 ? __ldg(a.q_cidx + q) : 0u;
          ++q;
        } else {
          rep = alpha ^ t_x[t];
          if (a.mode == kModeInversion) {  // BatchedOperator.chpl:187-199
            uint64_t const inverted = rep ^ a.inversion_mask;
            if (inverted < rep) { rep = inverted; c = 1; }
          }
        }
        double2 const ch = chars[c];
        double2 const v = t_v[t];
        double fr = ch.x * v.x - ch.y * v.y;  // Re(chi v): real vectors take the real part (DistributedMatrixVector.chpl:124)
        if (__popcll(alpha & t_s[t]) & 1) fr = -fr;
        double const coeff = fr * scale;
        if (coeff == 0.0) continue;  // :127 `if c != 0`
        int const d = push_owner(split, p.world, rep);
        unsigned const pos = atomicAdd(&counters[d], 1u);
        if (SCATTER) {
          PushRecord rec;
          rec.rep = rep;
          rec.c = coeff;
          p.send[pos] = rec;
        }
      }
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (int d = threadIdx.x; d < p.world; d += blockDim.x) p.hist[(size_t)d * p.nblocks + blockIdx.x] = counters[d];
  }
}

// displs[d] = start of owner d's records in the send buffer, displs[world] = their total
__global__ void push_displs_kernel(uint32_t const *__restrict__ base, int world, int nblocks,
                                   unsigned long long *__restrict__ displs) {
  int const d = threadIdx.x;
  if (d <= world) displs[d] = base[(size_t)d * nblocks];
}

// y[r] = sum_t [alpha_r & m_t == r_t] Re(v_t) sign_t x[r]  (kernels/reference.c:67-95): also what zero-initialises y
__global__ void __launch_bounds__(256)
push_diag_kernel(TermsView diag, uint64_t const *__restrict__ rows, int64_t n, double const *__restrict__ x,
                 double *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const alpha = rows[i];
    double d = 0.0;
    for (int t = 0; t < diag.number_terms; ++t)
      if ((alpha & __ldg(diag.m + t)) == __ldg(diag.r + t)) {
        double const v = __ldg(&diag.v[t].x);
        d += (__popcll(alpha & __ldg(diag.s + t)) & 1) ? -v : v;
      }
    y[i] = d * x[i];
  }
}

// Thread per received record: rank the representative among the LOCAL rows, y[j] += n_j c.
template <class Low>
__global__ void __launch_bounds__(256, 4)
push_consume_kernel(IndexView ix, GroupView g, int check_norm, double const *__restrict__ norms,
                    PushRecord const *__restrict__ records, int64_t count, double *__restrict__ y, int *error_flag) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
    ulonglong2 const raw = __ldcs(reinterpret_cast<ulonglong2 const *>(records) + k);
    uint64_t const needle[1] = {raw.x};
    double const c = __longlong_as_double((long long)raw.y);
    bool const live[1] = {true};
    int64_t j[1];
    if constexpr (std::is_void<Low>::value) index_find<1>(ix, needle, live, j);
    else index_find32<Low, 1>(ix, needle, live, j);
    if (j[0] >= 0) {
      atomicAdd(y + j[0], norms != nullptr ? __ldg(norms + j[0]) * c : c);
    } else {
      // not a row of this rank: fine when the state's norm vanishes, the reference's "invalid index" halt otherwise
      // (DistributedMatrixVector.chpl:127-135)
      bool bad = true;
      if (check_norm) bad = stabiliser_sum_global(g, raw.x) > kNormThreshold;
      if (bad) atomicOr(error_flag, 1);
    }
  }
}

struct PushSetup {
  MatvecArgs a{};
  int np = 4;
  bool inv = false;
  size_t count_smem = 0, orbit_smem = 0, emit_smem = 0;
  int64_t capacity = 0;
};

static void ensure_error_flag(MatvecScratch &sc) {
  Runtime &rt = runtime();
  if (sc.d_error == nullptr) {
    CUDA_CHECK(cudaMalloc(&sc.d_error, sizeof(int)));
    CUDA_CHECK(cudaMemsetAsync(sc.d_error, 0, sizeof(int), rt.stream));
    double const plain[6] = {1.0, 0.0, 1.0, 0.0, -1.0, 0.0};
    CUDA_CHECK(cudaMalloc(&sc.d_plain_chars, sizeof plain));
    CUDA_CHECK(cudaMemcpy(sc.d_plain_chars, plain, sizeof plain, cudaMemcpyHostToDevice));
  }
}

// The forward use of the term tables by the canonicalisation kernels: they match on `l` (adjoint form), so hand
// them a view with l and r exchanged.
static TermsView forward_view(TermsView t) {
  std::swap(t.l, t.r);
  return t;
}

static PushSetup push_setup(ls_hs_operator const *op, IndexData const &local) {
  Runtime &rt = runtime();
  MatvecScratch &sc = mv_scratch();
  ensure_error_flag(sc);
  ls_hs_basis const *basis = op->basis;
  OperatorDev &od = operator_dev(op);
  BasisInfo const info = basis_info(basis);
  PushSetup su;
  MatvecArgs &a = su.a;
  a.ix = local.view();
  a.rows = local.d_reps;
  a.off = od.off.view();
  a.diag = od.diag.view();
  a.error_flag = sc.d_error;
  a.spin_inversion = basis->spin_inversion;
  a.inversion_mask = basis->number_sites >= 64 ? ~uint64_t(0) : ((uint64_t(1) << basis->number_sites) - 1);
  a.mode = kModeNone;
  a.cvals = sc.d_plain_chars;
  a.number_chars = 1;
  a.number_vectors = 1;
  if (info.has_permutation_symmetries) {
    GroupData const &g = *info.group;
    LSB_CHECK(local.d_norms != nullptr || local.number_states == 0, "the local rows carry no norms");
    a.g = g.view();
    a.norms = local.d_norms;
    su.np = std::max(4, (g.number_bits + 3) / 4 * 4);
    su.inv = g.spin_inversion != 0;
    LSB_CHECK(!g.cinfo.empty(), "symmetry sectors with more than 256 distinct character values are not supported");
    a.cvals = g.d_cvals;
    a.number_chars = (int)(g.cvals.size() / 2);
    while ((1 << a.number_idx_planes) < a.number_chars) ++a.number_idx_planes;
    char const *env = getenv("LS_B200_MATVEC");
    bool const want_scalar = env != nullptr && strcmp(env, "scalar") == 0;
    bool const bitsliced = !want_scalar && orbit_prepare(g, su.np);
    a.mode = bitsliced ? kModeGroup : kModeGroupScalar;
    if (!bitsliced) a.number_idx_planes = std::max(a.number_idx_planes, 1);
  } else if (info.has_spin_inversion) {
    a.mode = kModeInversion;
    a.cvals = sc.d_plain_chars + (basis->spin_inversion < 0 ? 1 : 0);
    a.number_chars = 2;
  }
  int const T = a.off.number_terms;
  LSB_CHECK(T < 0x8000, "too many off-diagonal terms");
  su.count_smem = AdjointTerms::bytes(T, false);
  su.orbit_smem = ((su.count_smem + 15) & ~size_t(15)) + (size_t)(kOrbitThreads / 32) * kWarpSlabBytes;
  su.capacity = int64_t(1) << 26;
  if (char const *env = getenv("LS_B200_MV_CHUNK")) su.capacity = std::max<int64_t>(4096, atoll(env));
  LSB_CHECK(su.orbit_smem <= rt.smem_optin, "operator / symmetry tables do not fit in shared memory");
  return su;
}

// Rows per chunk: a function of the operator alone, so that every rank cuts its columns the same way and all ranks
// agree on the number of exchange rounds.
int64_t push_chunk_rows(ls_hs_operator const *op, int64_t local_rows) {
  OperatorDev &od = operator_dev(op);
  int64_t capacity = int64_t(1) << 26;
  if (char const *env = getenv("LS_B200_MV_CHUNK")) capacity = std::max<int64_t>(4096, atoll(env));
  int const T = std::max(1, od.off.number_terms);
  (void)local_rows;
  return std::max<int64_t>(1, capacity / T);
}

void push_begin(ls_hs_operator const *op, IndexData const &local, double const *d_x, double *d_y) {
  Runtime &rt = runtime();
  OperatorDev &od = operator_dev(op);
  int64_t const n = local.number_states;
  if (n == 0) return;
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 16);
  push_diag_kernel<<<blocks, 256, 0, rt.stream>>>(od.diag.view(), local.d_reps, n, d_x, d_y);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

void push_produce(ls_hs_operator const *op, IndexData const &local, DistShard &shard, int64_t chunk_begin,
                  int64_t chunk_rows, double const *d_x) {
  Runtime &rt = runtime();
  MatvecScratch &sc = mv_scratch();
  LSB_CHECK(shard.push != nullptr && shard.world <= kPushMaxWorld, "push buffers missing / too many ranks");
  PushBuffers &pb = *shard.push;
  int const world = shard.world;
  unsigned long long *displs = pb.displs.reserve((size_t)world + 1);
  if (chunk_rows <= 0) {
    CUDA_CHECK(cudaMemsetAsync(displs, 0, sizeof(unsigned long long) * ((size_t)world + 1), rt.stream));
    return;
  }
  PushSetup su = push_setup(op, local);
  MatvecArgs &a = su.a;
  int const T = a.off.number_terms;
  bool const queued = (a.mode == kModeGroup || a.mode == kModeGroupScalar) && T > 0;
  a.chunk_begin = chunk_begin;
  a.chunk_rows = (int)chunk_rows;
  ChunkSlot &slot = sc.slot[0];
  size_t const capacity = (size_t)chunk_rows * (size_t)std::max(T, 1);
  if (queued) {
    a.counts = slot.counts.reserve((size_t)chunk_rows + 1);
    a.offsets = slot.offsets.reserve((size_t)chunk_rows + 1);
    a.q_rep = slot.q_rep.reserve(capacity + 32);
    a.q_cidx = slot.q_cidx.reserve(capacity + 32);
    a.q_tsign = nullptr;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(chunk_rows + 1), rt.stream);
    if (tmp > sc.scan_tmp_bytes) {
      sc.scan_tmp.reserve(tmp);
      sc.scan_tmp_bytes = sc.scan_tmp.capacity;
    }
    // canonicalise with the FORWARD terms (the kernels match on `l`)
    MatvecArgs fa = a;
    fa.off = forward_view(a.off);
    allow_dynamic_smem(row_count_kernel, su.count_smem);
    row_count_kernel<<<ceil_div((size_t)chunk_rows + 1, 256), 256, su.count_smem, rt.stream>>>(fa);
    tmp = sc.scan_tmp_bytes;
    cub::DeviceScan::ExclusiveSum(sc.scan_tmp.ptr, tmp, fa.counts, fa.offsets, (int)(chunk_rows + 1), rt.stream);
    count_launch(2);
    if (a.mode == kModeGroup) {
      size_t const max_words = ((capacity + 1023) / 1024) * 32;
      orbit_launch(su.np, su.inv, false, max_words, su.orbit_smem, rt.stream, fa);
    } else {
      allow_dynamic_smem(orbit_scalar_kernel, su.count_smem);
      orbit_scalar_kernel<<<ceil_div((size_t)chunk_rows, 128), 128, su.count_smem, rt.stream>>>(fa);
    }
    count_launch();
    CUDA_CHECK(cudaGetLastError());
  }
  // group the records by owner: count -> scan -> scatter
  int const nblocks = (int)ceil_div((size_t)chunk_rows, kPushThreads);
  size_t const cells = (size_t)world * (size_t)nblocks + 1;
  PushArgs p{};
  p.a = a;
  p.x = d_x;
  p.splitters = shard.d_splitters;
  p.world = world;
  p.nblocks = nblocks;
  p.hist = pb.hist.reserve(cells);
  uint32_t *base = pb.base.reserve(cells);
  p.base = base;
  p.send = pb.send.reserve(capacity + 1);
  size_t const emit_smem = (size_t)T * 48 + (size_t)a.number_chars * 16 + (size_t)world * 12 + 16;
  LSB_CHECK(emit_smem <= rt.smem_optin, "operator tables do not fit in shared memory");
  allow_dynamic_smem(push_emit_kernel<false>, emit_smem);
  allow_dynamic_smem(push_emit_kernel<true>, emit_smem);
  CUDA_CHECK(cudaMemsetAsync(p.hist + cells - 1, 0, sizeof(uint32_t), rt.stream));
  push_emit_kernel<false><<<nblocks, kPushThreads, emit_smem, rt.stream>>>(p);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, p.hist, base, (int)cells, rt.stream);
  void *d_tmp = pb.scan_tmp.reserve(tmp);
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp, p.hist, base, (int)cells, rt.stream);
  push_emit_kernel<true><<<nblocks, kPushThreads, emit_smem, rt.stream>>>(p);
  push_displs_kernel<<<1, kPushMaxWorld + 32, 0, rt.stream>>>(base, world, nblocks, displs);
  count_launch(4);
  CUDA_CHECK(cudaGetLastError());
}

void push_consume(ls_hs_operator const *op, IndexData const &local, PushRecord const *records, int64_t count,
                  double *d_y) {
  if (count <= 0) return;
  Runtime &rt = runtime();
  MatvecScratch &sc = mv_scratch();
  ensure_error_flag(sc);
  BasisInfo const info = basis_info(op->basis);
  GroupView g{};
  int check_norm = 0;
  if (info.has_permutation_symmetries) {
    g = info.group->view();
    check_norm = 1;
  }
  IndexView const ix = local.view();
  unsigned const blocks = (unsigned)std::min<int64_t>((count + 255) / 256, (int64_t)rt.sm_count * 8);
  auto kernel = push_consume_kernel<void>;
  if (ix.offsets32 != nullptr && !ix.identity)
    kernel = ix.lows16 != nullptr   ? push_consume_kernel<uint16_t>
             : ix.lows32 != nullptr ? push_consume_kernel<uint32_t>
                                    : push_consume_kernel<uint64_t>;
  kernel<<<blocks, 256, 0, rt.stream>>>(ix, g, check_norm, local.d_norms, records, count, d_y, sc.d_error);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

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
  // num_vectors > 1 is an extension (the reference halts, DistributedMatrixVector.chpl:1096-1097):
  // x and y hold num_vectors contiguous vectors of length dim; all share one canonicalisation pass.
  if (num_vectors < 1) {
    ls_hs_error("matrix_vector_product: the number of vectors must be positive");
    return;
  }
  bool ok = true;
  guarded(__func__, [&] {
    IndexData *ix = index_of(op->basis);
    LSB_CHECK(ix != nullptr, "basis is not built");
    int64_t const dim = ix->number_states;
    MatvecScratch &sc = mv_scratch();
    cudaStream_t s = runtime().stream;
    size_t const n = (size_t)dim * (size_t)num_vectors;
    if (ix->dist != nullptr) {
      // The reference's per-locale block of the distributed product (DistributedMatrixVector.chpl:1060-1088): x and y
      // are THIS rank's blocks of the vectors; every rank of the communicator makes the same call.
      LSB_CHECK(num_vectors == 1, "distributed products take one vector at a time");
      double *d_x = sc.x.reserve(n + 1);
      double *d_y = sc.y.reserve(n + 1);
      cudaPointerAttributes attr_x{}, attr_y{};
      bool const pinned_x = cudaPointerGetAttributes(&attr_x, x) == cudaSuccess && attr_x.type == cudaMemoryTypeHost;
      bool const pinned_y = cudaPointerGetAttributes(&attr_y, y) == cudaSuccess && attr_y.type == cudaMemoryTypeHost;
      (void)cudaGetLastError();
      if (pinned_x && pinned_y && n > 0) {
        // x uploads on the copy stream under the canonicalisation of the first chunk; finished row chunks of y drain
        // on the copy stream behind the next chunks' kernels
        if (sc.copy_stream == nullptr) {
          CUDA_CHECK(cudaStreamCreateWithFlags(&sc.copy_stream, cudaStreamNonBlocking));
          CUDA_CHECK(cudaEventCreateWithFlags(&sc.copies_done, cudaEventDisableTiming));
        }
        if (sc.x_uploaded == nullptr) CUDA_CHECK(cudaEventCreateWithFlags(&sc.x_uploaded, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(sc.copies_done, s));  // (the previous product is through with d_x)
        CUDA_CHECK(cudaStreamWaitEvent(sc.copy_stream, sc.copies_done, 0));
        CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, sc.copy_stream));
        CUDA_CHECK(cudaEventRecord(sc.x_uploaded, sc.copy_stream));
        dist_matvec_local(op, d_x, d_y, 0, false, sc.x_uploaded, y);
      } else {
        if (n > 0) CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        dist_matvec_local(op, d_x, d_y, 0, false, nullptr, nullptr);
        if (n > 0) CUDA_CHECK(cudaMemcpyAsync(y, d_y, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
      }
      ok = matvec_finish();
      return;
    }
    if (dim == 0) return;
    double *d_x = sc.x.reserve(n);
    double *d_y = sc.y.reserve(n);
    // pinned y: finished row chunks drain on a copy stream behind the next chunks' kernels; pageable y (where an
    // "async" copy blocks the host and would stall the launch loop): one copy at the end
    cudaPointerAttributes attr{};
    bool const pinned = cudaPointerGetAttributes(&attr, y) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    // One vector from pinned memory: the upload of x runs on the copy stream under the canonicalise phase (which
    // does not read x); needs the canonicalised elements of ALL chunks at once (11 B each), so only below 24 GB.
    OperatorDev &od = operator_dev(op);
    if (od.stats_index != ix || od.stats_rows != dim) {
      od.stats_elements = count_elements(od, ix->d_reps, 0, dim);
      od.stats_index = ix;
      od.stats_rows = dim;
    }
    bool two_phases = pinned && num_vectors == 1 && od.stats_elements > 0 &&
                      od.stats_elements * 11 <= (int64_t(24) << 30) && getenv("LS_B200_NO_PHASED_E2E") == nullptr;
    if (two_phases && !(sc.phase_op == op && sc.phase_index == ix && !sc.phase_slots.empty())) {
      // First phased product of this operator: the persistent buffers (11 B per element, +25 % growth slack of
      // DeviceBuffer::reserve, + the 8 B values of the largest chunk) must fit next to everything already resident.
      size_t free_bytes = 0, total_bytes = 0;
      CUDA_CHECK(cudaMemGetInfo(&free_bytes, &total_bytes));
      free_bytes += block_cache_idle_bytes();
      size_t const needed = (size_t)((double)od.stats_elements * 11.0 * 1.25) + (size_t(3) << 30);
      if (needed > free_bytes) two_phases = false;
    }
    if (two_phases) {
      MatvecScratch &scr = mv_scratch();
      if (scr.copy_stream == nullptr) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&scr.copy_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&scr.copies_done, cudaEventDisableTiming));
      }
      if (scr.x_uploaded == nullptr) CUDA_CHECK(cudaEventCreateWithFlags(&scr.x_uploaded, cudaEventDisableTiming));
      CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, scr.copy_stream));
      CUDA_CHECK(cudaEventRecord(scr.x_uploaded, scr.copy_stream));
      // an allocation failure inside the phased path is not fatal: release its buffers and take the chunked path
      try {
        matvec_device(op, 0, dim, nullptr, nullptr, false, 1, 0, 0, nullptr, 1);
      } catch (CudaFailure const &) {
        (void)cudaGetLastError();
        release_phase_slots();
        two_phases = false;
      }
      CUDA_CHECK(cudaStreamWaitEvent(s, scr.x_uploaded, 0));
      if (two_phases) matvec_device(op, 0, dim, d_x, d_y, false, 1, 0, 0, y, 2);
      else matvec_device(op, 0, dim, d_x, d_y, false, 1, dim, dim, y);
    } else {
      CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, s));
      matvec_device(op, 0, dim, d_x, d_y, false, num_vectors, dim, dim, pinned ? y : nullptr);
    }
    if (!pinned) CUDA_CHECK(cudaMemcpyAsync(y, d_y, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
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

int ls_b200_matvec_device_phase(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, void const *x_dev,
                                void *y_dev, int complex_vectors, int phase) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(phase == 1 || phase == 2, "phase must be 1 (canonicalise) or 2 (apply)");
    matvec_device(op, row_begin, row_end, static_cast<double const *>(x_dev), static_cast<double *>(y_dev),
                  complex_vectors != 0, 1, 0, 0, nullptr, phase);
    status = 0;
  });
  return status;
}

int ls_b200_matvec_block_device(ls_hs_operator const *op, int64_t row_begin, int64_t row_end, int number_vectors,
                                double const *x_dev, int64_t x_stride, double *y_dev, int64_t y_stride) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(number_vectors >= 1, "the number of vectors must be positive");
    matvec_device(op, row_begin, row_end, x_dev, y_dev, false, number_vectors, x_stride, y_stride);
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

void ls_b200_operator_release(ls_hs_operator const *op) {
  if (op == nullptr) return;
  // Called from finalizers (Python's garbage collector): never wait for a library call in flight on another thread --
  // that call may itself be waiting for the interpreter (error handler).  Skipping only leaks a few small tables.
  std::unique_lock<std::mutex> lock(runtime().mutex, std::try_to_lock);
  if (!lock.owns_lock()) return;
  MatvecScratch &sc = mv_scratch();
  if (sc.phase_op == op) release_phase_slots();
  operator_cache().erase(op);
}

int64_t ls_b200_count_matrix_elements(ls_hs_operator const *op, int64_t row_begin, int64_t row_end) {
  int64_t n = -1;
  guarded(__func__, [&] {
    IndexData *ix = index_of(op->basis);
    LSB_CHECK(ix != nullptr, "basis is not built");
    n = count_elements(operator_dev(op), ix->d_reps, row_begin, row_end);
  });
  return n;
}

}  // extern "C"
