// bitslice.cuh -- bit-sliced ("32 states per register") orbit arithmetic.
//
// The reference walks every group element's Benes network per state
// (kernels/generator.cpp:5-8, 88-94: depth x 6 u64 ops per element).  On the
// GPU we transpose 32 states into bit planes -- plane i holds bit i of 32
// states -- so that a permutation of bits is a *renaming of planes* (a shared
// memory address, zero ALU work) and the only arithmetic left is a bit-serial
// lexicographic comparison: two LOP3 per plane per group element for 32
// states.  Outputs are identical to the reference's; only the schedule differs.
//
// Everything here is __host__ __device__ so that tests/ can run the exact
// same code on the CPU against the oracle.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define LSB_HD __host__ __device__ __forceinline__
#else
#define LSB_HD inline
#endif

namespace lsb {

// In-place transpose of a 32x32 bit matrix held as 32 words:
// afterwards a[i] bit k == (before) a[k] bit i.
LSB_HD void transpose32(uint32_t (&a)[32]) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if ((k & j) == 0) {
        // swap the high-j bits of a[k] with the low-j bits of a[k+j]
        uint32_t const t = ((a[k] >> j) ^ a[k + j]) & m;
        a[k + j] ^= t;
        a[k] ^= t << j;
      }
    }
  }
}

// 32 ascending states whose bits 16..31 all agree (consecutive fixed-weight candidates almost always do):
// planes 16..31 are constants and planes 0..15 come from a 16x16 butterfly on packed halves (states k and
// k + 16 share a word) -- four stages on 16 words instead of five on 32.  In place: a[i] becomes plane i.
LSB_HD void transpose32_low16(uint32_t (&a)[32]) {
  uint32_t const upper = a[0] >> 16;
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = (a[k] & 0xffffu) | (a[k + 16] << 16);
  uint32_t m = 0x00ff00ffu;
#pragma unroll
  for (int j = 8; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if ((k & j) == 0) {
        uint32_t const t = ((w[k] >> j) ^ w[k + j]) & m;
        w[k + j] ^= t;
        w[k] ^= t << j;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = w[i];
    a[i + 16] = ((upper >> i) & 1u) ? 0xffffffffu : 0u;
  }
}

// Bit-serial comparison state for 32 lanes, fed least-significant plane first.
//   lt: y < x so far,  eq: y == x so far.
// One LOP3 each: a higher differing plane overrides the verdict of lower ones.
LSB_HD void cmp_step(uint32_t y, uint32_t x, uint32_t &lt, uint32_t &eq) {
  uint32_t const d = y ^ x;
  lt = (d & x) | (~d & lt);
  eq &= ~d;
}
// Same for the spin-flipped image ~y (flip restricted to the live planes).
LSB_HD void cmp_step_flipped(uint32_t y, uint32_t x, uint32_t &lt, uint32_t &eq) {
  uint32_t const d = y ^ x;  // (~y) ^ x == ~d
  lt = (~d & x) | (d & lt);
  eq &= d;
}

// Running-minimum update for state_info: where lt is set the image replaces
// the current minimum plane.
LSB_HD uint32_t select_plane(uint32_t lt, uint32_t y, uint32_t r) { return (lt & y) | (~lt & r); }

}  // namespace lsb
