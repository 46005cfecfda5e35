// bitslice_harness.cpp -- the bit-sliced orbit arithmetic of the CUDA kernels, run on the HOST.
//
// csrc/bitslice.cuh is __host__ __device__: this file includes THAT header (not a copy) and drives its primitives
// (transpose32, transpose32_low16, cmp_step, cmp_step_flipped, select_plane) through the same group loop the kernels
// run -- csrc/orbit.cu (running minimum + minimising element: state_info) and csrc/basis_build.cu
// (alive / event bits: is_representative) -- with the shared-memory column replaced by an array and the constant-memory
// plane table by the permutations themselves.  tests/test_bitslice_cpu.py compares the results with the oracle, so the
// arithmetic idea (planes renamed instead of bits permuted, min(y, ~y) = y ^ top(y) for spin inversion, the
// least-significant-plane-first comparison) is checked on the CPU tier, for every plane count up to 64.
//
// Built by the test with g++ (no CUDA involved):  g++ -O2 -shared -fPIC -o tests/_build/libbitslice_harness.so ...
#include <cstdint>
#include <cstring>

#include "../lattice_symmetries_b200/csrc/bitslice.cuh"

using namespace lsb;

namespace {

// 32 states -> planes[0 .. 63] (plane i = bit i of the 32 states); `low16` selects the shortcut the build kernel takes
// when the upper halves of the low words agree (basis_build.cu:344).
void to_planes(uint64_t const *states, uint32_t (&planes)[64], bool try_low16) {
  uint32_t lo[32], hi[32];
  for (int k = 0; k < 32; ++k) {
    lo[k] = (uint32_t)states[k];
    hi[k] = (uint32_t)(states[k] >> 32);
  }
  bool const same_hi = hi[0] == hi[31];
  bool ascending = true;
  for (int k = 1; k < 32; ++k) ascending = ascending && states[k] >= states[k - 1];
  if (try_low16 && ascending && same_hi && ((lo[0] ^ lo[31]) >> 16) == 0) transpose32_low16(lo);
  else transpose32(lo);
  transpose32(hi);
  for (int i = 0; i < 32; ++i) {
    planes[i] = lo[i];
    planes[32 + i] = hi[i];
  }
}

void from_planes(uint32_t const (&planes)[64], uint64_t *states) {
  uint32_t lo[32], hi[32];
  for (int i = 0; i < 32; ++i) {
    lo[i] = planes[i];
    hi[i] = planes[32 + i];
  }
  transpose32(lo);
  transpose32(hi);
  for (int k = 0; k < 32; ++k) states[k] = ((uint64_t)hi[k] << 32) | lo[k];
}

}  // namespace

extern "C" {

// transpose32 against the definition, its involution, and transpose32_low16 against transpose32 under its precondition.
int bitslice_selftest(uint64_t seed) {
  auto next = [&seed]() {
    seed += 0x9E3779B97F4A7C15ull;
    uint64_t z = seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  for (int trial = 0; trial < 200; ++trial) {
    uint32_t a[32], b[32];
    for (int k = 0; k < 32; ++k) a[k] = b[k] = (uint32_t)next();
    transpose32(b);
    for (int i = 0; i < 32; ++i)
      for (int k = 0; k < 32; ++k)
        if (((b[i] >> k) & 1u) != ((a[k] >> i) & 1u)) return 1;
    transpose32(b);
    if (memcmp(a, b, sizeof a) != 0) return 2;
    // low16: 32 words whose bits 16..31 agree
    uint32_t const upper = (uint32_t)next() & 0xffff0000u;
    uint32_t c[32], d[32];
    for (int k = 0; k < 32; ++k) c[k] = d[k] = upper | ((uint32_t)next() & 0xffffu);
    transpose32(c);
    transpose32_low16(d);
    if (memcmp(c, d, sizeof c) != 0) return 3;
  }
  return 0;
}

// state_info, bit-sliced (csrc/orbit.cu:118-163): for every state the orbit minimum, the first element that reaches
// it (-1: the state itself, i.e. the initial value of the running tuple, generator.cpp:105-106) and whether it was the
// spin-flipped image.  perm[j * nbits + i] = source bit of output bit i under element j (Benes.hs:338-345).
void bitslice_state_info(int nbits, int G, int32_t const *perm, int inversion, int64_t number_words,
                         uint64_t const *states, uint64_t *representatives, int32_t *argmin, uint8_t *flipped) {
  for (int64_t w = 0; w < number_words; ++w) {
    uint32_t planes[64];
    to_planes(states + 32 * w, planes, false);
    uint32_t r[64];
    memcpy(r, planes, sizeof r);
    int32_t *arg = argmin + 32 * w;
    uint8_t *flip = flipped + 32 * w;
    for (int k = 0; k < 32; ++k) {
      arg[k] = -1;
      flip[k] = 0;
    }
    for (int j = 0; j < G; ++j) {
      int32_t const *p = perm + (size_t)j * nbits;
      // z = min(y, ~y) = y ^ top(y): the flipped image is the smaller one iff the top live bit of y is set
      uint32_t const top = inversion != 0 ? planes[p[nbits - 1]] : 0u;
      uint32_t z[64];
      uint32_t lt = 0, eq = 0xffffffffu;
      for (int i = 0; i < 64; ++i) {
        z[i] = i < nbits ? (planes[p[i]] ^ top) : 0u;
        cmp_step(z[i], r[i], lt, eq);  // least significant plane first: a higher plane overrides
      }
      for (int i = 0; i < 64; ++i) r[i] = select_plane(lt, z[i], r[i]);
      for (int k = 0; k < 32; ++k)
        if ((lt >> k) & 1u) {
          arg[k] = j;
          flip[k] = (uint8_t)((top >> k) & 1u);
        }
    }
    from_planes(r, representatives + 32 * w);
  }
}

// is_representative, bit-sliced (csrc/basis_build.cu:366-399): alive = no image (nor flipped image) is smaller than
// the state; events = some image other than element `skip` (the identity, or -1) EQUALS the state -- directly or
// flipped -- i.e. the stabiliser is not trivial and the exact character sum has to decide the norm.
// Returns the number of (word, element) pairs in which the one-comparison form disagreed with comparing y and ~y
// separately (must be 0).
int64_t bitslice_is_representative(int nbits, int G, int32_t const *perm, int inversion, int skip, int64_t number_words,
                                   uint64_t const *states, uint8_t *alive_out, uint8_t *events_out, int try_low16) {
  int64_t disagreements = 0;
  for (int64_t w = 0; w < number_words; ++w) {
    uint32_t planes[64];
    to_planes(states + 32 * w, planes, try_low16 != 0);
    uint32_t alive = 0xffffffffu, events = 0;
    for (int j = 0; j < G; ++j) {
      int32_t const *p = perm + (size_t)j * nbits;
      uint32_t const top = inversion != 0 ? planes[p[nbits - 1]] : 0u;
      uint32_t lt = 0, eq = 0xffffffffu;      // z = min(y, ~y) against x
      uint32_t lt_f = 0, eq_f = 0xffffffffu;  // the flipped image ~y against x, through cmp_step_flipped
      for (int i = 0; i < 64; ++i) {
        uint32_t const y = i < nbits ? planes[p[i]] : 0u;
        cmp_step(i < nbits ? (y ^ top) : 0u, planes[i], lt, eq);
        if (i < nbits) cmp_step_flipped(y, planes[i], lt_f, eq_f);
        else cmp_step(0u, planes[i], lt_f, eq_f);
      }
      alive &= ~lt;
      // min(y, ~y) < x <=> y < x or ~y < x, always; min(y, ~y) == x <=> y == x or ~y == x for every x whose top live
      // bit is clear (min(y, ~y) never has it set) -- the only x the build enumerates under spin inversion
      // (Basis.hs:736-740), and the only ones that can be alive.  Checked against the two separate comparisons.
      if (inversion != 0) {
        uint32_t lt_y = 0, eq_y = 0xffffffffu;
        for (int i = 0; i < 64; ++i) cmp_step(i < nbits ? planes[p[i]] : 0u, planes[i], lt_y, eq_y);
        uint32_t const x_top = planes[nbits - 1];
        if ((lt_y | lt_f) != lt || (((eq_y | eq_f) ^ eq) & ~x_top) != 0) ++disagreements;
      }
      if (j != skip) events |= eq;
    }
    events &= alive;
    for (int k = 0; k < 32; ++k) {
      alive_out[32 * w + k] = (uint8_t)((alive >> k) & 1u);
      events_out[32 * w + k] = (uint8_t)((events >> k) & 1u);
    }
  }
  return disagreements;
}

}  // extern "C"
