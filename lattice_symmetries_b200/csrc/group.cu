// group.cu -- symmetry-group tables on the device and the two per-state batch
// kernels behind the reference's Halide symbols:
//   ls_internal_create/destroy_halide_kernel_data   kernels/kernels.c:112-193
//   ls_hs_state_info_halide_kernel                  kernels/kernels.c:246-319
//   ls_hs_is_representative_halide_kernel           kernels/kernels.c:195-244
// and the dispatch wrappers of kernels/reference.c:137-171.
#include <atomic>

#include "state.hpp"

namespace lsb {

GroupView GroupData::view() const {
  GroupView v{};
  v.number_bits = number_bits;
  v.depth = depth;
  v.number_masks = number_masks;
  v.spin_inversion = spin_inversion;
  v.flip_mask = flip_mask;
  v.masks = d_masks;
  v.re = d_re;
  v.im = d_im;
  v.perm = d_perm;
  v.cvals = d_cvals;
  v.cinfo = d_cinfo;
  v.number_chars = cinfo.empty() ? 0 : (int)(cvals.size() / 2);
  for (int k = 0; k < depth && k < kMaxDepth; ++k) v.shifts[k] = (unsigned)shifts[k];
  return v;
}

GroupData::~GroupData() {
  cudaFree(d_masks);
  cudaFree(d_re);
  cudaFree(d_im);
  cudaFree(d_perm);
  cudaFree(d_cvals);
  cudaFree(d_cinfo);
  magic = 0;
}

// ---- kernels ------------------------------------------------------------------
template <class W>
__global__ void __launch_bounds__(256)
state_info_kernel(GroupView g, int64_t n, uint64_t const *__restrict__ alphas,
                  uint64_t *__restrict__ betas, double2 *__restrict__ characters,
                  double *__restrict__ norms) {
  extern __shared__ unsigned char smem_raw[];
  W *smasks = reinterpret_cast<W *>(smem_raw);
  stage_masks<W>(g, smasks);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t rep;
    double c_re, c_im, acc;
    state_info_scalar<W>(g, smasks, alphas[i], rep, c_re, c_im, acc);
    betas[i] = rep;
    characters[i] = make_double2(c_re, c_im);
    norms[i] = norm_from_sum(g, acc);
  }
}

// kernels/generator.cpp:181-253: flag = all images >= x, norm = raw stabiliser
// character sum; flag forced to 0 unless the sum is positive.  We stop at the
// first image below x (the reference stops per SIMD chunk, :236), so the sum is
// only meaningful where the flag is set -- as in the reference.
template <class W>
__global__ void __launch_bounds__(256)
is_representative_kernel(GroupView g, int64_t n, uint64_t const *__restrict__ alphas,
                         uint8_t *__restrict__ flags, double *__restrict__ norms) {
  extern __shared__ unsigned char smem_raw[];
  W *smasks = reinterpret_cast<W *>(smem_raw);
  stage_masks<W>(g, smasks);
  OrbitScalar<W> orbit{g, smasks};
  W const flip = (W)g.flip_mask;
  int const inv = g.spin_inversion;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    W const x = (W)alphas[i];
    bool flag = true;
    double s = 0.0;
#pragma unroll 1
    for (int j = 0; j < g.number_masks; ++j) {
      W const y = orbit.image(x, j);
      if (y < x) { flag = false; break; }
      if (y == x) s += g.re[j];
      if (inv != 0) {
        W const yf = y ^ flip;
        if (yf < x) { flag = false; break; }
        if (yf == x) s += (double)inv * g.re[j];
      }
    }
    norms[i] = s;
    flags[i] = (flag && s > kNormThreshold) ? 1 : 0;
  }
}

static size_t masks_smem_bytes(GroupData const &g, bool narrow) {
  return (size_t)g.depth * (size_t)g.number_masks * (narrow ? 4 : 8);
}

template <class K>
static void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void launch_state_info(GroupData const &g, int64_t n, uint64_t const *d_alphas, uint64_t *d_betas,
                       double2 *d_chars, double *d_norms) {
  if (n == 0) return;
  Runtime &rt = runtime();
  bool const narrow = g.number_bits <= 32;
  size_t const smem = masks_smem_bytes(g, narrow);
  LSB_CHECK(smem <= rt.smem_optin, "symmetry group too large for shared memory staging");
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
  if (narrow) {
    allow_smem(state_info_kernel<uint32_t>, smem);
    state_info_kernel<uint32_t><<<blocks, 256, smem, rt.stream>>>(g.view(), n, d_alphas, d_betas, d_chars, d_norms);
  } else {
    allow_smem(state_info_kernel<uint64_t>, smem);
    state_info_kernel<uint64_t><<<blocks, 256, smem, rt.stream>>>(g.view(), n, d_alphas, d_betas, d_chars, d_norms);
  }
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

void launch_is_representative(GroupData const &g, int64_t n, uint64_t const *d_alphas,
                              uint8_t *d_flags, double *d_norms) {
  if (n == 0) return;
  Runtime &rt = runtime();
  bool const narrow = g.number_bits <= 32;
  size_t const smem = masks_smem_bytes(g, narrow);
  LSB_CHECK(smem <= rt.smem_optin, "symmetry group too large for shared memory staging");
  unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
  if (narrow) {
    allow_smem(is_representative_kernel<uint32_t>, smem);
    is_representative_kernel<uint32_t><<<blocks, 256, smem, rt.stream>>>(g.view(), n, d_alphas, d_flags, d_norms);
  } else {
    allow_smem(is_representative_kernel<uint64_t>, smem);
    is_representative_kernel<uint64_t><<<blocks, 256, smem, rt.stream>>>(g.view(), n, d_alphas, d_flags, d_norms);
  }
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

// Scratch for the host-pointer batch entry points.
struct BatchScratch {
  DeviceBuffer<uint64_t> alphas, betas;
  DeviceBuffer<double2> chars;
  DeviceBuffer<double> norms;
  DeviceBuffer<uint8_t> flags;
  std::vector<uint64_t> gather;
};
static BatchScratch &scratch() {
  static BatchScratch s;
  return s;
}
constexpr int64_t kBatchChunk = int64_t(1) << 22;

}  // namespace lsb

using namespace lsb;

extern "C" {

void *ls_internal_create_halide_kernel_data(ls_hs_permutation_group const *g, int const spin_inversion) {
  if (spin_inversion != 0 && spin_inversion != 1 && spin_inversion != -1)
    ls_hs_fatal_error(__func__, __LINE__, "invalid spin_inversion");  // kernels.c:186-187
  static std::atomic<uint64_t> next_id{1};
  auto *self = new GroupData();
  self->id = next_id.fetch_add(1);
  self->number_bits = g->number_bits;
  self->depth = g->number_shifts;
  self->number_masks = g->number_masks;
  self->spin_inversion = spin_inversion;
  // kernels.c:96-98 get_flip_mask_64
  self->flip_mask = g->number_bits == 0 ? 0 : (~uint64_t(0) >> (64 - g->number_bits));
  LSB_CHECK(self->depth <= kMaxDepth, "Benes network too deep");
  size_t const G = (size_t)g->number_masks, D = (size_t)g->number_shifts;
  if (G * D > 0) self->masks.assign(g->masks, g->masks + G * D);
  if (D > 0) self->shifts.assign(g->shifts, g->shifts + D);
  if (G > 0) {
    self->re.assign(g->eigvals_re, g->eigvals_re + G);
    self->im.assign(g->eigvals_im, g->eigvals_im + G);
  }
  for (double v : self->im)
    if (v != 0.0) self->real_characters = false;
  // Recover the site permutations from the networks: the image of the state
  // with only bit j set has exactly the bits i with perm[i] == j set.
  int const nb = g->number_bits;
  self->perm.assign(G * (size_t)nb, 0);
  for (size_t e = 0; e < G; ++e)
    for (int j = 0; j < nb; ++j) {
      uint64_t y = uint64_t(1) << j;
      for (size_t k = 0; k < D; ++k) {
        uint64_t const m = self->masks[k * G + e], d = self->shifts[k];
        uint64_t const t = ((y >> d) ^ y) & m;
        y = (y ^ t) ^ (t << d);
      }
      LSB_CHECK(y != 0 && (y & (y - 1)) == 0 && (nb == 64 || (y >> nb) == 0),
                "Benes network is not a permutation of the live bits");
      self->perm[e * (size_t)nb + (size_t)__builtin_ctzll(y)] = (uint8_t)j;
    }
  // Distinct character values: the bit-sliced matvec tracks, per state, the
  // index of the character of the minimising image instead of the element.
  {
    std::vector<double> &cv = self->cvals;
    cv = {1.0, 0.0};
    auto find = [&cv](double re, double im) -> int {
      for (size_t k = 0; k < cv.size() / 2; ++k)
        if (memcmp(&cv[2 * k], &re, 8) == 0 && memcmp(&cv[2 * k + 1], &im, 8) == 0) return (int)k;
      cv.push_back(re);
      cv.push_back(im);
      return (int)(cv.size() / 2 - 1);
    };
    self->cinfo.resize(G);
    bool fits = true;
    for (size_t e = 0; e < G && fits; ++e) {
      int const a = find(self->re[e], self->im[e]);
      int b = 0;
      if (spin_inversion != 0) b = find((double)spin_inversion * self->re[e], (double)spin_inversion * self->im[e]);
      if (a > 255 || b > 255) fits = false;
      self->cinfo[e] = (uint16_t)(a | (b << 8));
    }
    if (!fits) self->cinfo.clear();
  }
  guarded(__func__, [&] {
    if (G == 0) return;
    CUDA_CHECK(cudaMalloc(&self->d_cvals, sizeof(double) * self->cvals.size()));
    CUDA_CHECK(cudaMemcpy(self->d_cvals, self->cvals.data(), sizeof(double) * self->cvals.size(), cudaMemcpyHostToDevice));
    if (!self->cinfo.empty()) {
      CUDA_CHECK(cudaMalloc(&self->d_cinfo, sizeof(uint16_t) * G));
      CUDA_CHECK(cudaMemcpy(self->d_cinfo, self->cinfo.data(), sizeof(uint16_t) * G, cudaMemcpyHostToDevice));
    }
    CUDA_CHECK(cudaMalloc(&self->d_masks, sizeof(uint64_t) * G * std::max<size_t>(D, 1)));
    CUDA_CHECK(cudaMalloc(&self->d_re, sizeof(double) * G));
    CUDA_CHECK(cudaMalloc(&self->d_im, sizeof(double) * G));
    CUDA_CHECK(cudaMalloc(&self->d_perm, std::max<size_t>(G * (size_t)nb, 1)));
    cudaStream_t s = runtime().stream;
    if (D > 0) CUDA_CHECK(cudaMemcpyAsync(self->d_masks, self->masks.data(), sizeof(uint64_t) * G * D, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(self->d_re, self->re.data(), sizeof(double) * G, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(self->d_im, self->im.data(), sizeof(double) * G, cudaMemcpyHostToDevice, s));
    if (nb > 0) CUDA_CHECK(cudaMemcpyAsync(self->d_perm, self->perm.data(), G * (size_t)nb, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
  return self;
}

void ls_internal_destroy_halide_kernel_data(void *p) {
  if (p == nullptr) return;
  auto *self = static_cast<GroupData *>(p);
  LSB_CHECK(self->magic == kGroupMagic, "not a kernel-data object of this library");
  std::lock_guard<std::mutex> lock(runtime().mutex);
  delete self;
}

void ls_hs_state_info_halide_kernel(ptrdiff_t batch_size, uint64_t const *alphas,
                                    ptrdiff_t alphas_stride, uint64_t *betas,
                                    ptrdiff_t betas_stride, ls_hs_scalar *characters,
                                    double *norms, void const *private_data) {
  auto const *g = static_cast<GroupData const *>(private_data);
  LSB_CHECK(g != nullptr && g->magic == kGroupMagic, "invalid state_info kernel data");
  guarded(__func__, [&] {
    BatchScratch &sc = scratch();
    cudaStream_t s = runtime().stream;
    for (int64_t begin = 0; begin < batch_size; begin += kBatchChunk) {
      int64_t const n = std::min<int64_t>(kBatchChunk, batch_size - begin);
      uint64_t const *src = alphas + begin * alphas_stride;
      if (alphas_stride != 1) {
        sc.gather.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) sc.gather[(size_t)i] = src[i * alphas_stride];
        src = sc.gather.data();
      }
      uint64_t *d_a = sc.alphas.reserve((size_t)n);
      uint64_t *d_b = sc.betas.reserve((size_t)n);
      double2 *d_c = sc.chars.reserve((size_t)n);
      double *d_n = sc.norms.reserve((size_t)n);
      CUDA_CHECK(cudaMemcpyAsync(d_a, src, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
      launch_state_info(*g, n, d_a, d_b, d_c, d_n);
      if (betas_stride == 1) {
        CUDA_CHECK(cudaMemcpyAsync(betas + begin, d_b, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
      } else {
        sc.gather.resize((size_t)n);
        CUDA_CHECK(cudaMemcpyAsync(sc.gather.data(), d_b, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
      }
      CUDA_CHECK(cudaMemcpyAsync(characters + begin, d_c, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaMemcpyAsync(norms + begin, d_n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (betas_stride != 1)
        for (int64_t i = 0; i < n; ++i) betas[(begin + i) * betas_stride] = sc.gather[(size_t)i];
    }
  });
}

void ls_hs_is_representative_halide_kernel(ptrdiff_t batch_size, uint64_t const *alphas,
                                           ptrdiff_t alphas_stride, uint8_t *are_representatives,
                                           double *norms, void const *private_data) {
  auto const *g = static_cast<GroupData const *>(private_data);
  LSB_CHECK(g != nullptr && g->magic == kGroupMagic, "invalid is_representative kernel data");
  guarded(__func__, [&] {
    BatchScratch &sc = scratch();
    cudaStream_t s = runtime().stream;
    for (int64_t begin = 0; begin < batch_size; begin += kBatchChunk) {
      int64_t const n = std::min<int64_t>(kBatchChunk, batch_size - begin);
      uint64_t const *src = alphas + begin * alphas_stride;
      if (alphas_stride != 1) {
        sc.gather.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) sc.gather[(size_t)i] = src[i * alphas_stride];
        src = sc.gather.data();
      }
      uint64_t *d_a = sc.alphas.reserve((size_t)n);
      uint8_t *d_f = sc.flags.reserve((size_t)n);
      double *d_n = sc.norms.reserve((size_t)n);
      CUDA_CHECK(cudaMemcpyAsync(d_a, src, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
      launch_is_representative(*g, n, d_a, d_f, d_n);
      CUDA_CHECK(cudaMemcpyAsync(are_representatives + begin, d_f, (size_t)n, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaMemcpyAsync(norms + begin, d_n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
  });
}

// ---- kernels/reference.c:137-171 dispatch wrappers ------------------------------
void ls_hs_state_index(ls_hs_basis const *basis, ptrdiff_t batch_size, uint64_t const *spins,
                       ptrdiff_t spins_stride, ptrdiff_t *indices, ptrdiff_t indices_stride) {
  LSB_CHECK(basis->kernels->state_index_kernel != nullptr, "state_index_kernel is NULL");
  (*basis->kernels->state_index_kernel)(batch_size, spins, spins_stride, indices, indices_stride,
                                        basis->kernels->state_index_data);
}

void ls_hs_is_representative(ls_hs_basis const *basis, ptrdiff_t batch_size, uint64_t const *alphas,
                             ptrdiff_t alphas_stride, uint8_t *are_representatives, double *norms) {
  LSB_CHECK(basis->kernels->is_representative_kernel != nullptr,
            "is_representative_kernel is NULL, perhaps this basis requires no projection?");
  (*basis->kernels->is_representative_kernel)(batch_size, alphas, alphas_stride, are_representatives,
                                              norms, basis->kernels->is_representative_data);
}

void ls_hs_state_info(ls_hs_basis const *basis, ptrdiff_t batch_size, uint64_t const *alphas,
                      ptrdiff_t alphas_stride, uint64_t *betas, ptrdiff_t betas_stride,
                      ls_hs_scalar *characters, double *norms) {
  LSB_CHECK(basis->kernels->state_info_kernel != nullptr,
            "state_info_kernel is NULL, perhaps this basis requires no projection?");
  (*basis->kernels->state_info_kernel)(batch_size, alphas, alphas_stride, betas, betas_stride,
                                       characters, norms, basis->kernels->state_info_data);
}

}  // extern "C"
