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

}  // extern "C"
