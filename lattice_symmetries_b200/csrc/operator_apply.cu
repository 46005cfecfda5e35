// operator_apply.cu -- "rows of H" queries: apply the operator's diagonal /
// off-diagonal terms to a batch of basis states.
//
// Replaces kernels/reference.c:67-134 (ls_internal_operator_apply_diag_x1,
// ls_internal_operator_apply_off_diag_x1) and the Chapel exports built on them,
// chapel/src/BatchedOperator.chpl:298-357.  Host pointers in, host arrays out,
// exactly like the reference; the work runs on the device as count -> scan ->
// fill so that the emitted (beta, coeff) pairs keep the reference's order
// (states in input order, terms in table order).
#include <cub/device/device_scan.cuh>

#include "state.hpp"

namespace lsb {

__global__ void __launch_bounds__(256)
apply_diag_kernel(TermsView diag, int64_t n, uint64_t const *__restrict__ alphas,
                  double const *__restrict__ xs, double *__restrict__ ys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const alpha = alphas[i];
    double acc = 0.0;
    for (int t = 0; t < diag.number_terms; ++t)
      if ((alpha & __ldg(diag.m + t)) == __ldg(diag.r + t)) {
        int const sign = 1 - 2 * (__popcll(alpha & __ldg(diag.s + t)) & 1);
        double const factor = xs != nullptr ? sign * xs[i] : (double)sign;
        acc += __ldg(&diag.v[t].x) * factor;  // creal(v), reference.c:89
      }
    ys[i] = acc;
  }
}

__global__ void __launch_bounds__(256)
off_diag_count_kernel(TermsView off, int64_t n, uint64_t const *__restrict__ alphas,
                      int64_t *__restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = 0;
    if (i < n) {
      uint64_t const alpha = alphas[i];
      for (int t = 0; t < off.number_terms; ++t) c += ((alpha & __ldg(off.m + t)) == __ldg(off.r + t)) ? 1 : 0;
    }
    counts[i] = c;  // counts[n] = 0 so that the exclusive scan yields offsets[n] = total
  }
}

__global__ void __launch_bounds__(256)
off_diag_fill_kernel(TermsView off, int64_t n, uint64_t const *__restrict__ alphas,
                     double const *__restrict__ xs, int64_t const *__restrict__ offsets,
                     uint64_t *__restrict__ betas, double2 *__restrict__ coeffs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const alpha = alphas[i];
    int64_t q = offsets[i];
    double const scale = xs != nullptr ? xs[i] : 1.0;
    for (int t = 0; t < off.number_terms; ++t)
      if ((alpha & __ldg(off.m + t)) == __ldg(off.r + t)) {
        int const sign = 1 - 2 * (__popcll(alpha & __ldg(off.s + t)) & 1);
        double const factor = sign * scale;
        double2 const v = __ldg(off.v + t);
        betas[q] = alpha ^ __ldg(off.x + t);
        coeffs[q] = make_double2(v.x * factor, v.y * factor);
        ++q;
      }
  }
}

struct ApplyScratch {
  DeviceBuffer<uint64_t> alphas, betas;
  DeviceBuffer<double> xs, ys;
  DeviceBuffer<double2> coeffs;
  DeviceBuffer<int64_t> counts, offsets;
  DeviceBuffer<unsigned char> scan_tmp;
};
static ApplyScratch &apply_scratch() {
  static ApplyScratch s;
  return s;
}

static unsigned grid_for(int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)runtime().sm_count * 16));
}

// Returns the number of emitted pairs; outputs are host pointers (betas /
// coeffs may be nullptr when only the offsets are wanted).
static int64_t apply_off_diag(ls_hs_operator const *op, int64_t n, uint64_t const *alphas, uint64_t *betas,
                              ls_hs_scalar *coeffs, ptrdiff_t *offsets, double const *xs) {
  Runtime &rt = runtime();
  ApplyScratch &sc = apply_scratch();
  OperatorDev &od = operator_dev(op);
  cudaStream_t s = rt.stream;
  if (od.off.number_terms == 0 || n == 0) {
    for (int64_t i = 0; i <= n; ++i) offsets[i] = 0;
    return 0;
  }
  uint64_t *d_a = sc.alphas.reserve((size_t)n);
  int64_t *d_counts = sc.counts.reserve((size_t)n + 1);
  int64_t *d_offsets = sc.offsets.reserve((size_t)n + 1);
  double *d_xs = nullptr;
  CUDA_CHECK(cudaMemcpyAsync(d_a, alphas, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
  if (xs != nullptr) {
    d_xs = sc.xs.reserve((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(d_xs, xs, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  }
  off_diag_count_kernel<<<grid_for(n + 1), 256, 0, s>>>(od.off.view(), n, d_a, d_counts);
  count_launch();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_offsets, n + 1, s);
  unsigned char *tmp = sc.scan_tmp.reserve(tmp_bytes);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_counts, d_offsets, n + 1, s);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  static_assert(sizeof(ptrdiff_t) == sizeof(int64_t), "ptrdiff_t must be 64-bit");
  CUDA_CHECK(cudaMemcpyAsync(offsets, d_offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  int64_t const total = offsets[n];
  if (total > 0 && betas != nullptr) {
    uint64_t *d_b = sc.betas.reserve((size_t)total);
    double2 *d_c = sc.coeffs.reserve((size_t)total);
    off_diag_fill_kernel<<<grid_for(n), 256, 0, s>>>(od.off.view(), n, d_a, d_xs, d_offsets, d_b, d_c);
    count_launch();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(betas, d_b, sizeof(uint64_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(coeffs, d_c, sizeof(double2) * (size_t)total, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return total;
}

static void apply_diag(ls_hs_operator const *op, int64_t n, uint64_t const *alphas, double *ys, double const *xs) {
  Runtime &rt = runtime();
  ApplyScratch &sc = apply_scratch();
  OperatorDev &od = operator_dev(op);
  cudaStream_t s = rt.stream;
  if (n == 0) return;
  if (od.diag.number_terms == 0) {  // reference.c:73-76
    memset(ys, 0, (size_t)n * sizeof(double));
    return;
  }
  uint64_t *d_a = sc.alphas.reserve((size_t)n);
  double *d_y = sc.ys.reserve((size_t)n);
  double *d_xs = nullptr;
  CUDA_CHECK(cudaMemcpyAsync(d_a, alphas, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
  if (xs != nullptr) {
    d_xs = sc.xs.reserve((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(d_xs, xs, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  }
  apply_diag_kernel<<<grid_for(n), 256, 0, s>>>(od.diag.view(), n, d_a, d_xs, d_y);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(ys, d_y, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

// ---- projected rows of H (semantics of chapel/src/BatchedOperator.chpl:207-253, 264-282) ----------
// For every alpha_i: H|alpha_i> = sum_k c_k |beta_k>; every beta is canonicalised -- representative, character, norm
// (state_info) -- and the coefficient becomes chi_k c_k n(beta_k) / n(alpha_i): the matrix element
// <rep_k| H |alpha_i> of the symmetry-projected operator.  Zero-norm images keep their slot with coefficient 0, so
// the offsets are those of the un-projected call.
__global__ void __launch_bounds__(256)
project_rows_kernel(int64_t n, int64_t const *__restrict__ offsets, double2 const *__restrict__ chars,
                    double const *__restrict__ norms_beta, double const *__restrict__ norms_alpha,
                    double2 *__restrict__ coeffs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double const na = norms_alpha[i];
    for (int64_t k = offsets[i]; k < offsets[i + 1]; ++k) {
      double const nb = norms_beta[k];
      double2 const c = coeffs[k];
      // same order as the oracle / the reference: (c n_beta / n_alpha) first, then times the character
      double const tr = c.x * nb / na, ti = c.y * nb / na;
      double2 const ch = chars[k];
      coeffs[k] = nb > 0.0 ? make_double2(ch.x * tr - ch.y * ti, ch.x * ti + ch.y * tr) : make_double2(0.0, 0.0);
    }
  }
}
// spin inversion without permutations (BatchedOperator.chpl:187-199): rep = min(beta, ~beta), coefficient * character
__global__ void __launch_bounds__(256)
project_inversion_kernel(int64_t total, uint64_t mask, double character, uint64_t *__restrict__ betas,
                         double2 *__restrict__ coeffs) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    uint64_t const b = betas[k], f = b ^ mask;
    if (f < b) {
      betas[k] = f;
      coeffs[k] = make_double2(coeffs[k].x * character, coeffs[k].y * character);
    }
  }
}

void launch_state_info(GroupData const &g, int64_t n, uint64_t const *d_alphas, uint64_t *d_betas, double2 *d_chars,
                       double *d_norms);
void launch_state_index(IndexData const &ix, int64_t n, uint64_t const *d_needles, int64_t *d_out);

static int64_t apply_off_diag_projected(ls_hs_operator const *op, int64_t n, uint64_t const *alphas, uint64_t *reps,
                                        ls_hs_scalar *coeffs, int64_t *offsets, int64_t *indices) {
  Runtime &rt = runtime();
  ApplyScratch &sc = apply_scratch();
  OperatorDev &od = operator_dev(op);
  cudaStream_t s = rt.stream;
  if (od.off.number_terms == 0 || n == 0) {
    for (int64_t i = 0; i <= n; ++i) offsets[i] = 0;
    return 0;
  }
  BasisInfo const info = basis_info(op->basis);
  uint64_t *d_a = sc.alphas.reserve((size_t)n);
  int64_t *d_counts = sc.counts.reserve((size_t)n + 1);
  int64_t *d_offsets = sc.offsets.reserve((size_t)n + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_a, alphas, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
  off_diag_count_kernel<<<grid_for(n + 1), 256, 0, s>>>(od.off.view(), n, d_a, d_counts);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_offsets, n + 1, s);
  unsigned char *tmp = sc.scan_tmp.reserve(tmp_bytes);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_counts, d_offsets, n + 1, s);
  count_launch(2);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(offsets, d_offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  int64_t const total = offsets[n];
  if (total == 0) return 0;
  uint64_t *d_b = sc.betas.reserve((size_t)total);
  double2 *d_c = sc.coeffs.reserve((size_t)total);
  off_diag_fill_kernel<<<grid_for(n), 256, 0, s>>>(od.off.view(), n, d_a, nullptr, d_offsets, d_b, d_c);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
  static DeviceBuffer<uint64_t> rep_buffer, alpha_reps;
  static DeviceBuffer<double2> char_buffer, alpha_chars;
  static DeviceBuffer<double> norm_beta, norm_alpha;
  static DeviceBuffer<int64_t> index_buffer;
  uint64_t *d_out = d_b;
  if (info.has_permutation_symmetries) {
    GroupData const &g = *info.group;
    d_out = rep_buffer.reserve((size_t)total);
    double2 *d_ch = char_buffer.reserve((size_t)total);
    double *d_nb = norm_beta.reserve((size_t)total);
    double *d_na = norm_alpha.reserve((size_t)n);
    launch_state_info(g, total, d_b, d_out, d_ch, d_nb);
    // norms of the alphas themselves (any states are allowed, not only this basis' representatives)
    launch_state_info(g, n, d_a, alpha_reps.reserve((size_t)n), alpha_chars.reserve((size_t)n), d_na);
    project_rows_kernel<<<grid_for(n), 256, 0, s>>>(n, d_offsets, d_ch, d_nb, d_na, d_c);
    count_launch();
  } else if (info.has_spin_inversion) {
    uint64_t const mask = op->basis->number_sites >= 64 ? ~uint64_t(0) : ((uint64_t(1) << op->basis->number_sites) - 1);
    project_inversion_kernel<<<grid_for(total), 256, 0, s>>>(total, mask, (double)op->basis->spin_inversion, d_b, d_c);
    count_launch();
  }
  CUDA_CHECK(cudaGetLastError());
  if (indices != nullptr) {
    IndexData const *ix = index_of(op->basis);
    LSB_CHECK(ix != nullptr, "indices were asked for, but the basis is not built");
    int64_t *d_j = index_buffer.reserve((size_t)total);
    launch_state_index(*ix, total, d_out, d_j);
    CUDA_CHECK(cudaMemcpyAsync(indices, d_j, sizeof(int64_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  CUDA_CHECK(cudaMemcpyAsync(reps, d_out, sizeof(uint64_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaMemcpyAsync(coeffs, d_c, sizeof(double2) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  return total;
}

static void free_host(void *p) { free(p); }

template <class T>
static chpl_external_array make_external(size_t count) {
  chpl_external_array arr;
  arr.elts = count > 0 ? malloc(count * sizeof(T)) : nullptr;
  arr.num_elts = count;
  arr.freer = count > 0 ? reinterpret_cast<void *>(&free_host) : nullptr;
  return arr;
}

}  // namespace lsb

using namespace lsb;

extern "C" {

// kernels/reference.c:67-95
void ls_internal_operator_apply_diag_x1(ls_hs_operator const *op, ptrdiff_t batch_size, uint64_t const *alphas,
                                        double *ys, double const *xs) {
  guarded(__func__, [&] { apply_diag(op, batch_size, alphas, ys, xs); });
}

// kernels/reference.c:97-134
void ls_internal_operator_apply_off_diag_x1(ls_hs_operator const *op, ptrdiff_t batch_size,
                                            uint64_t const *alphas, uint64_t *betas, ls_hs_scalar *coeffs,
                                            ptrdiff_t *offsets, double const *xs) {
  guarded(__func__, [&] { apply_off_diag(op, batch_size, alphas, betas, coeffs, offsets, xs); });
}

// chapel/src/BatchedOperator.chpl:298-316
void ls_chpl_operator_apply_diag(ls_hs_operator *op, int64_t count, uint64_t *alphas, chpl_external_array *coeffs,
                                 int64_t num_tasks) {
  (void)num_tasks;
  if (op->basis->requires_projection) {
    ls_hs_error("bases that require projection are not yet supported");  // :307-308
    return;
  }
  *coeffs = make_external<double>((size_t)count);
  guarded(__func__, [&] { apply_diag(op, count, alphas, static_cast<double *>(coeffs->elts), nullptr); });
}

// chapel/src/BatchedOperator.chpl:318-357.  The reference sizes betas / coeffs
// as count * numberOffDiagTerms and leaves the tail unused; so do we.
void ls_chpl_operator_apply_off_diag(ls_hs_operator *op, int64_t count, uint64_t *alphas,
                                     chpl_external_array *betas, chpl_external_array *coeffs,
                                     chpl_external_array *offsets, int64_t num_tasks) {
  (void)num_tasks;
  int const T = op->off_diag_terms != nullptr ? op->off_diag_terms->number_terms : 0;
  *offsets = make_external<int64_t>((size_t)count + 1);
  if (T == 0) {
    *betas = chpl_external_array{nullptr, 0, nullptr};
    *coeffs = chpl_external_array{nullptr, 0, nullptr};
    for (int64_t i = 0; i <= count; ++i) static_cast<int64_t *>(offsets->elts)[i] = 0;
    return;
  }
  *betas = make_external<uint64_t>((size_t)count * (size_t)T);
  *coeffs = make_external<ls_hs_scalar>((size_t)count * (size_t)T);
  guarded(__func__, [&] {
    apply_off_diag(op, count, alphas, static_cast<uint64_t *>(betas->elts),
                   static_cast<ls_hs_scalar *>(coeffs->elts), static_cast<ptrdiff_t *>(offsets->elts), nullptr);
  });
}

// Extension: the rows of the PROJECTED operator, batched.  Host pointers; reps / coeffs need room for
// count * (number of off-diagonal terms) entries, offsets for count + 1; indices (optional, same length as reps)
// receives the position of every representative in the built basis, -1 when absent.  Returns the number of entries, -1
// on error.
int64_t ls_b200_operator_apply_off_diag_projected(ls_hs_operator const *op, int64_t count, uint64_t const *alphas,
                                                  uint64_t *reps, ls_hs_scalar *coeffs, int64_t *offsets,
                                                  int64_t *indices) {
  int64_t total = -1;
  guarded(__func__, [&] { total = apply_off_diag_projected(op, count, alphas, reps, coeffs, offsets, indices); });
  return total;
}

}  // extern "C"
