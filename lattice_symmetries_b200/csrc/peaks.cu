// peaks.cu -- measured integer-issue ceiling of the device the library runs on.
//
// Both hot kernels are bound by 32-bit logic-op issue (LOP3 / XOR on the ALU
// pipe), not by HBM.  MEASURED_PEAKS.json carries only HBM and
// tensor peaks, so the roofline denominator for the integer side is measured
// here, live, by a register-resident LOP3 chain: 8 independent accumulators per
// thread, enough resident warps to saturate every SM sub-partition.
#include "state.hpp"

namespace lsb {

__global__ void __launch_bounds__(256)
lop3_peak_kernel(uint32_t *__restrict__ sink, uint32_t seed, int iterations) {
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u;
  uint32_t a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
#pragma unroll 1
  for (int i = 0; i < iterations; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      // one LOP3 (bitwise select, LUT 0xCA) each; 8 chains, each link depends on three live registers
      a0 = (a0 & a1) | (~a0 & a2);
      a1 = (a1 & a2) | (~a1 & a3);
      a2 = (a2 & a3) | (~a2 & a4);
      a3 = (a3 & a4) | (~a3 & a5);
      a4 = (a4 & a5) | (~a4 & a6);
      a5 = (a5 & a6) | (~a5 & a7);
      a6 = (a6 & a7) | (~a6 & a0);
      a7 = (a7 & a0) | (~a7 & a1);
    }
  }
  uint32_t const r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (r == 0x12345678u) sink[blockIdx.x] = r;  // practically never; keeps the chain alive
}

}  // namespace lsb

using namespace lsb;

extern "C" double ls_b200_measure_lop3_peak(void) {
  double result = 0.0;
  guarded(__func__, [&] {
    Runtime &rt = runtime();
    uint32_t *sink = nullptr;
    int const blocks = rt.sm_count * 8, iterations = 4096;
    CUDA_CHECK(cudaMalloc(&sink, sizeof(uint32_t) * (size_t)blocks));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CUDA_CHECK(cudaEventRecord(e0, rt.stream));
      lop3_peak_kernel<<<blocks, 256, 0, rt.stream>>>(sink, 12345u + (uint32_t)rep, iterations);
      count_launch();
      CUDA_CHECK(cudaEventRecord(e1, rt.stream));
      CUDA_CHECK(cudaStreamSynchronize(rt.stream));
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    double const ops = (double)blocks * 256.0 * (double)iterations * 16.0 * 8.0;
    result = ops / ((double)best * 1e-3);
  });
  return result;
}
