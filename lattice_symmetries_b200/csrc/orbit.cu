// orbit.cu -- canonicalisation of matrix elements: the bit-sliced orbit walk (integer-issue bound).
//
// Replaces the per-element ls_hs_state_info calls of chapel/src/BatchedOperator.chpl:207-253
// (kernels/generator.cpp:24-54, 77-140 arithmetic): one thread canonicalises 32 matrix elements at once,
// see bitslice.cuh / plane_table.cuh.  Kept in its own translation unit: the 16 x 2 x 2 template
// instances take minutes to compile and own the constant-memory plane table.
#include "matvec_args.cuh"
#include "plane_table.cuh"

namespace lsb {

// One thread canonicalises 32 consecutive matrix elements of the chunk; a warp
// owns 1024 consecutive elements and a private 8 KB slab of shared memory that
// first stages its betas and then holds its bit planes.
template <int NP, bool INV>
__global__ void __launch_bounds__(kOrbitThreads, (NP <= 40 ? 5 : NP <= 52 ? 4 : 3))  // NP <= 40: five CTAs per SM (<= 102 registers)
orbit_kernel(MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  bool const pack_tsign = NP <= 48 && a.q_tsign != nullptr;
  AdjointTerms terms;
  terms.stage(smem, a.off, pack_tsign);
  size_t const terms_bytes = (AdjointTerms::bytes(a.off.number_terms, pack_tsign) + 15) & ~size_t(15);
  __syncthreads();

  int const tid = threadIdx.x;
  int const lane = tid & 31;
  unsigned char *slab = smem + terms_bytes + (size_t)(tid >> 5) * kWarpSlabBytes;
  uint64_t *stage = reinterpret_cast<uint64_t *>(slab);   // [k][lane]: element 32 * lane + k of the warp
  uint32_t *planes = reinterpret_cast<uint32_t *>(slab);  // [plane][lane], after the betas have been read
  uint32_t const total = a.offsets[a.chunk_rows];
  // One CTA per four 1024-element blocks.  (A persistent grid sized to the machine was measured 15 % slower on
  // kagome-36: the warps of an SM then walk through their load / integer phases in lockstep.)
  uint64_t const warp_q0 = ((uint64_t)blockIdx.x * (kOrbitThreads / 32) + (tid >> 5)) * 1024;
  if (warp_q0 >= total) return;  // whole warps leave: everything below runs converged
  uint64_t const warp_q1 = min((uint64_t)total, warp_q0 + 1024);
  uint64_t const q0 = warp_q0 + 32 * (uint64_t)lane;
  int const lanes = q0 < total ? (int)min((uint64_t)32, total - q0) : 0;

  // ---- regenerate the warp's betas: lane = row, CSR positions from the offsets -------
  {
    // first row with elements in [warp_q0, warp_q1): offsets[row] <= warp_q0 < offsets[row + 1]
    int lo_r = 0, hi_r = a.chunk_rows;
    while (hi_r - lo_r > 1) {
      int const mid = (lo_r + hi_r) >> 1;
      if (__ldg(a.offsets + mid) <= (uint32_t)warp_q0) lo_r = mid; else hi_r = mid;
    }
    int const T = terms.T;
    for (int base = lo_r;; base += 32) {
      int const row = base + lane;
      uint32_t q = row < a.chunk_rows ? __ldg(a.offsets + row) : 0xffffffffu;
      bool const mine = row < a.chunk_rows && q < warp_q1;
      if (!__any_sync(0xffffffffu, mine)) break;
      if (mine) {
        uint64_t const alpha = __ldg(a.rows + a.chunk_begin + row);
        for (int t = 0; t < T; ++t)
          if ((alpha & terms.m[t]) == terms.l[t]) {
            if (q >= warp_q0 && q < warp_q1) {
              unsigned const e = (unsigned)(q - warp_q0);
              uint64_t beta = alpha ^ terms.x[t];
              // the term and its sign ride in the 16 spare bits above the state (NP <= 48)
              if (pack_tsign)
                beta |= (uint64_t)((unsigned)t | ((unsigned)(__popcll(alpha & terms.s[t]) & 1) << 15)) << 48;
              stage[(e & 31u) * 32 + (e >> 5)] = beta;
            }
            ++q;
          }
      }
    }
  }
  __syncwarp();
  uint32_t r[NP];
  {
    uint32_t lo[32], hi[32];
    uint64_t beta = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < lanes) beta = stage[k * 32 + lane];
      lo[k] = (uint32_t)beta;  // lanes past the end repeat the last element (or carry zeros)
      hi[k] = (uint32_t)(beta >> 32);
    }
    __syncwarp();  // every lane holds its betas: the slab may now be overwritten by the planes
    if (pack_tsign) {
      uint16_t *ts_out = a.q_tsign + q0;
      if (lanes == 32) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 v;
          v.x = (hi[k] >> 16) | (hi[k + 1] & 0xffff0000u);
          v.y = (hi[k + 2] >> 16) | (hi[k + 3] & 0xffff0000u);
          v.z = (hi[k + 4] >> 16) | (hi[k + 5] & 0xffff0000u);
          v.w = (hi[k + 6] >> 16) | (hi[k + 7] & 0xffff0000u);
          *reinterpret_cast<uint4 *>(ts_out + k) = v;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < lanes) ts_out[k] = (uint16_t)(hi[k] >> 16);
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) hi[k] &= 0xffffu;
    }
    transpose32(lo);
    if (NP > 32) transpose32(hi);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      r[i] = (i < 32) ? lo[(i < 32) ? i : 0] : hi[(i < 32) ? 0 : i - 32];
      planes[i * 32 + lane] = r[i];
    }
  }
  __syncwarp();  // all 32 lanes are here (see above); a thread only reads back its own column
  unsigned char const *column = reinterpret_cast<unsigned char const *>(planes + lane);
  uint32_t idx[kMvIdxPlanes];
#pragma unroll
  for (int p = 0; p < kMvIdxPlanes; ++p) idx[p] = 0;  // character 0 == 1+0i: the input itself (generator.cpp:105-106)
  int const nbits = a.g.number_bits;
  int const nidx = a.number_idx_planes;
  int const G = (a.debug_skip & 1) ? 0 : a.g.number_masks;
#if LS_ORBIT_RETARGET
  uint32_t top = 0;
#endif
#pragma unroll 1
  for (int j = 0; j < G; ++j) {
    PlaneRow<NP> const po(j);
    // z = min(y, ~y) = y ^ top(y) when spin inversion is present (see basis_build.cu)
#if LS_ORBIT_RETARGET
    if (INV) {
      uint32_t const ro = po[NP + 2];
      if (ro != kNoRetarget) {
        uint32_t const d = *reinterpret_cast<uint32_t const *>(column + ro);
#pragma unroll
        for (int i = 0; i < NP; ++i)
          if (i < NP - 3 || i < nbits) planes[i * 32 + lane] ^= d;
        top ^= d;
      }
    }
#else
    uint32_t top = 0;
    if (INV) top = *reinterpret_cast<uint32_t const *>(column + po[NP]);
#endif
    uint32_t z[NP];
    uint32_t lt = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      z[i] = *reinterpret_cast<uint32_t const *>(column + po[i]);
#if !LS_ORBIT_RETARGET
      if (INV) z[i] ^= (i < NP - 3 || i < nbits) ? top : 0u;  // padding planes stay zero
#endif
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
  // back to one state per word, CSR order in global memory
  {
    uint32_t lo[32], hi[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      lo[i] = (i < NP) ? r[(i < NP) ? i : 0] : 0u;
      hi[i] = (i + 32 < NP) ? r[(i + 32 < NP) ? i + 32 : 0] : 0u;
    }
    transpose32(lo);
    if (NP > 32) transpose32(hi);
    uint64_t *out = a.q_rep + q0;
    if (lanes == 32) {
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        ulonglong2 v;
        v.x = (NP > 32) ? (((uint64_t)hi[k] << 32) | lo[k]) : (uint64_t)lo[k];
        v.y = (NP > 32) ? (((uint64_t)hi[k + 1] << 32) | lo[k + 1]) : (uint64_t)lo[k + 1];
        *reinterpret_cast<ulonglong2 *>(out + k) = v;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < lanes) out[k] = (NP > 32) ? (((uint64_t)hi[k] << 32) | lo[k]) : (uint64_t)lo[k];
    }
  }
  if (nidx > 0) {
    uint32_t info[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) info[i] = (i < kMvIdxPlanes) ? idx[(i < kMvIdxPlanes) ? i : 0] : 0u;
    transpose32(info);
    uint8_t *out = a.q_cidx + q0;
    if (lanes == 32) {
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<uint32_t *>(out + k) =
            (info[k] & 0xffu) | ((info[k + 1] & 0xffu) << 8) | ((info[k + 2] & 0xffu) << 16) | (info[k + 3] << 24);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < lanes) out[k] = (uint8_t)info[k];
    }
  }
}

// ---- fused path: canonicalise, rank and gather in one kernel ------------------------
// Same walk as orbit_kernel, but the representatives never leave the SM: after the
// group loop they are transposed back into the warp's slab and every lane ranks and
// gathers one element per step (kFusedBatch searches in flight per lane), so the
// memory latency of the lookups hides behind the integer work of the other warps.
// Output: one value per matrix element, CSR order, summed per row by row_sum_kernel
// (deterministic order, no atomics).


// Last phase of orbit_gather_kernel: every lane ranks kFusedBatch representatives at a
// time (independent searches in flight), gathers n_j x_j, applies conj(chi) w sign and
// stores the values in CSR order.  Low = void: the generic 64-bit index path.
struct FusedLookup {
  MatvecArgs const &a;
  unsigned char const *slab;
  uint16_t const *tsign;
  uint8_t const *cslab;
  double2 const *chars;
  double2 const *tw;
  uint64_t warp_q0;
  int count;
  int lane;

  template <class Low>
  __device__ __forceinline__ void run() const {
    uint64_t const *out = reinterpret_cast<uint64_t const *>(slab);
    bool const skip_gather = (a.debug_skip & 2) != 0;
    bool const cplx = a.complex_vectors != 0;
    int const nidx = a.number_idx_planes;
#pragma unroll 1
    for (int it = 0; it * 32 < count; it += kFusedBatch) {
      uint64_t needle[kFusedBatch];
      bool live[kFusedBatch];
#pragma unroll
      for (int u = 0; u < kFusedBatch; ++u) {
        live[u] = (it + u) * 32 + lane < count && !skip_gather;
        needle[u] = live[u] ? out[lane * kOutPitch + it + u] : 0;
      }
      int64_t j[kFusedBatch];
      if constexpr (std::is_void<Low>::value) index_find<kFusedBatch>(a.ix, needle, live, j);
      else index_find32<Low, kFusedBatch>(a.ix, needle, live, j);
      double2 xv[kFusedBatch];
#pragma unroll
      for (int u = 0; u < kFusedBatch; ++u) {
        xv[u] = make_double2(0.0, 0.0);
        if (j[u] >= 0) {
          if (cplx) xv[u] = __ldg(reinterpret_cast<double2 const *>(a.xs) + j[u]);
          else xv[u].x = __ldg(a.xs + j[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kFusedBatch; ++u) {
        int const e = (it + u) * 32 + lane;
        if (e >= count) continue;
        double vr = 0.0, vi = 0.0;
        if (live[u]) {
          unsigned const ts = tsign[e];
          double2 w = tw[ts & 0x7fffu];
          if (ts & 0x8000u) { w.x = -w.x; w.y = -w.y; }
          unsigned const c = nidx > 0 ? cslab[lane * kCidxPitch + it + u] : 0u;
          double2 const ch = chars[c];
          double const fr = ch.x * w.x + ch.y * w.y;  // conj(chi) * w
          double const fi = ch.x * w.y - ch.y * w.x;
          if (j[u] >= 0) {
            vr = fr * xv[u].x - fi * xv[u].y;
            vi = fr * xv[u].y + fi * xv[u].x;
          } else if (fr != 0.0 || fi != 0.0) {
            // not in the basis: fine when its norm vanishes, an error otherwise
            // (DistributedMatrixVector.chpl:127-135)
            if (stabiliser_sum_global(a.g, needle[u]) > kNormThreshold) atomicOr(a.error_flag, 1);
          }
        }
        if (cplx) reinterpret_cast<double2 *>(a.vals)[warp_q0 + e] = make_double2(vr, vi);
        else a.vals[warp_q0 + e] = vr;
      }
    }
  }
};

template <int NP, bool INV>
__global__ void __launch_bounds__(kOrbitThreads, 4)
orbit_gather_kernel(__grid_constant__ MatvecArgs const a) {
  extern __shared__ __align__(16) unsigned char smem[];
  AdjointTerms terms;
  terms.stage(smem, a.off, true);
  size_t const terms_bytes = (AdjointTerms::bytes(a.off.number_terms, true) + 15) & ~size_t(15);
  double2 *chars = reinterpret_cast<double2 *>(smem + terms_bytes);
  for (int j = threadIdx.x; j < a.number_chars; j += blockDim.x) chars[j] = a.cvals[j];
  __syncthreads();

  int const tid = threadIdx.x;
  int const lane = tid & 31;
  unsigned char *slab = smem + terms_bytes + (size_t)a.number_chars * 16 + (size_t)(tid >> 5) * kFusedWarpBytes;
  uint64_t *stage = reinterpret_cast<uint64_t *>(slab);   // [k][lane]: element 32 * lane + k of the warp
  uint32_t *planes = reinterpret_cast<uint32_t *>(slab);  // [plane][lane], after the betas have been read
  uint16_t *tsign = reinterpret_cast<uint16_t *>(slab + kFusedSlabBytes);
  uint8_t *cslab = slab + kFusedSlabBytes + kFusedTsignBytes;
  uint32_t const total = a.offsets[a.chunk_rows];
  uint64_t const warp_q0 = ((uint64_t)blockIdx.x * (kOrbitThreads / 32) + (tid >> 5)) * 1024;
  if (warp_q0 >= total) return;  // whole warps leave: everything below runs converged
  uint64_t const warp_q1 = min((uint64_t)total, warp_q0 + 1024);
  uint64_t const q0 = warp_q0 + 32 * (uint64_t)lane;
  int const lanes = q0 < total ? (int)min((uint64_t)32, total - q0) : 0;

  // ---- regenerate the warp's betas: lane = row, CSR positions from the offsets -------
  {
    int lo_r = 0, hi_r = a.chunk_rows;
    while (hi_r - lo_r > 1) {
      int const mid = (lo_r + hi_r) >> 1;
      if (__ldg(a.offsets + mid) <= (uint32_t)warp_q0) lo_r = mid; else hi_r = mid;
    }
    int const T = terms.T;
    for (int base = lo_r;; base += 32) {
      int const row = base + lane;
      uint32_t q = row < a.chunk_rows ? __ldg(a.offsets + row) : 0xffffffffu;
      bool const mine = row < a.chunk_rows && q < warp_q1;
      if (!__any_sync(0xffffffffu, mine)) break;
      if (mine) {
        uint64_t const alpha = __ldg(a.rows + a.chunk_begin + row);
        for (int t = 0; t < T; ++t)
          if ((alpha & terms.m[t]) == terms.l[t]) {
            if (q >= warp_q0 && q < warp_q1) {
              unsigned const e = (unsigned)(q - warp_q0);
              stage[(e & 31u) * 32 + (e >> 5)] = alpha ^ terms.x[t];
              tsign[e] = (uint16_t)((unsigned)t | ((unsigned)(__popcll(alpha & terms.s[t]) & 1) << 15));
            }
            ++q;
          }
      }
    }
  }
  __syncwarp();
  uint32_t r[NP];
  {
    uint32_t lo[32], hi[32];
    uint64_t beta = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < lanes) beta = stage[k * 32 + lane];
      lo[k] = (uint32_t)beta;  // lanes past the end repeat the last element (or carry zeros)
      hi[k] = (uint32_t)(beta >> 32);
    }
    __syncwarp();  // every lane holds its betas: the slab may now be overwritten by the planes
    transpose32(lo);
    if (NP > 32) transpose32(hi);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      r[i] = (i < 32) ? lo[(i < 32) ? i : 0] : hi[(i < 32) ? 0 : i - 32];
      planes[i * 32 + lane] = r[i];
    }
  }
  __syncwarp();
  unsigned char const *column = reinterpret_cast<unsigned char const *>(planes + lane);
  uint32_t idx[kMvIdxPlanes];
#pragma unroll
  for (int p = 0; p < kMvIdxPlanes; ++p) idx[p] = 0;
  int const nbits = a.g.number_bits;
  int const nidx = a.number_idx_planes;
  int const G = (a.debug_skip & 1) ? 0 : a.g.number_masks;
  uint32_t top = 0;  // flip plane of the current class of elements: bit (number_bits - 1) of their images
#pragma unroll 1
  for (int j = 0; j < G; ++j) {
    PlaneRow<NP> const po(j);
    // z = min(y, ~y) = y ^ top(y) when spin inversion is present (see basis_build.cu).  The
    // slab holds the planes already XOR-ed with the flip plane of the current class; when the
    // class changes (every |G| / number_bits rows) the copy is re-targeted in place.
    if (INV) {
      uint32_t const ro = po[NP + 2];
      if (ro != kNoRetarget) {
        uint32_t const d = *reinterpret_cast<uint32_t const *>(column + ro);
#pragma unroll
        for (int i = 0; i < NP; ++i)
          if (i < NP - 3 || i < nbits) planes[i * 32 + lane] ^= d;  // padding planes stay zero
        top ^= d;
      }
    }
    uint32_t z[NP];
    uint32_t lt = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      z[i] = *reinterpret_cast<uint32_t const *>(column + po[i]);
      lt = ((z[i] ^ r[i]) & r[i]) | (~(z[i] ^ r[i]) & lt);   // one LOP3: z < r, most significant plane last
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) r[i] = (lt & z[i]) | (~lt & r[i]);
    if (nidx > 0) {
      unsigned const ci = po[NP + 1];
      auto update = [&](int p) {
        uint32_t const c = 0u - ((ci >> p) & 1u);
        uint32_t const cf = 0u - ((ci >> (8 + p)) & 1u);
        uint32_t const nb = INV ? ((top & cf) | (~top & c)) : c;
        idx[p] = (lt & nb) | (~lt & idx[p]);
      };
      update(0);
      if (nidx > 1) update(1);
      if (nidx > 2) { update(2); update(3); }
      if (nidx > 4) { update(4); update(5); update(6); update(7); }
    }
  }
  __syncwarp();  // every lane is done with its plane column: the slab becomes the output slab
  // back to one state per word: out[k][owner lane] (pitch 33: conflict-free both ways)
  {
    uint32_t lo[32], hi[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      lo[i] = (i < NP) ? r[(i < NP) ? i : 0] : 0u;
      hi[i] = (i + 32 < NP) ? r[(i + 32 < NP) ? i + 32 : 0] : 0u;
    }
    transpose32(lo);
    if (NP > 32) transpose32(hi);
    uint64_t *out = reinterpret_cast<uint64_t *>(slab);
#pragma unroll
    for (int k = 0; k < 32; ++k)
      out[k * kOutPitch + lane] = (NP > 32) ? (((uint64_t)hi[k] << 32) | lo[k]) : (uint64_t)lo[k];
  }
  if (nidx > 0) {
    uint32_t info[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) info[i] = (i < kMvIdxPlanes) ? idx[(i < kMvIdxPlanes) ? i : 0] : 0u;
    transpose32(info);
#pragma unroll
    for (int k = 0; k < 32; ++k) cslab[k * kCidxPitch + lane] = (uint8_t)info[k];
  }
  __syncwarp();

  // ---- rank + gather: step `it` handles elements it * 32 + lane (coalesced stores) -----
  FusedLookup const L{a, slab, tsign, cslab, chars, terms.w, warp_q0, (int)(warp_q1 - warp_q0), lane};
  if (a.ix.offsets32 != nullptr && !a.ix.identity) {
    if (a.ix.lows16 != nullptr) L.template run<uint16_t>();
    else if (a.ix.lows32 != nullptr) L.template run<uint32_t>();
    else L.template run<uint64_t>();
  } else {
    L.template run<void>();
  }
}

using OrbitKernel = void (*)(MatvecArgs);
template <int NP>
static OrbitKernel pick_orbit_inv(bool inv, bool fused) {
  if (fused) return inv ? orbit_gather_kernel<NP, true> : orbit_gather_kernel<NP, false>;
  return inv ? orbit_kernel<NP, true> : orbit_kernel<NP, false>;
}
static OrbitKernel pick_orbit_kernel(int np, bool inv, bool fused) {
  switch (np) {
    case 4: return pick_orbit_inv<4>(inv, fused);
    case 8: return pick_orbit_inv<8>(inv, fused);
    case 12: return pick_orbit_inv<12>(inv, fused);
    case 16: return pick_orbit_inv<16>(inv, fused);
    case 20: return pick_orbit_inv<20>(inv, fused);
    case 24: return pick_orbit_inv<24>(inv, fused);
    case 28: return pick_orbit_inv<28>(inv, fused);
    case 32: return pick_orbit_inv<32>(inv, fused);
    case 36: return pick_orbit_inv<36>(inv, fused);
    case 40: return pick_orbit_inv<40>(inv, fused);
    case 44: return pick_orbit_inv<44>(inv, fused);
    case 48: return pick_orbit_inv<48>(inv, fused);
    case 52: return pick_orbit_inv<52>(inv, fused);
    case 56: return pick_orbit_inv<56>(inv, fused);
    case 60: return pick_orbit_inv<60>(inv, fused);
    case 64: return pick_orbit_inv<64>(inv, fused);
  }
  return nullptr;
}

bool orbit_prepare(GroupData const &g, int np) { return upload_plane_offsets(g, np, 32); }

void orbit_launch(int np, bool inv, bool fused, size_t words, size_t smem, cudaStream_t stream, MatvecArgs const &a) {
  OrbitKernel kernel = pick_orbit_kernel(np, inv, fused);
  LSB_CHECK(kernel != nullptr, "unsupported number of bits");
  if (smem > 48 * 1024) {
    // (per kernel instance; cheap, and idempotent)
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (words == 0) return;
  kernel<<<ceil_div(words, kOrbitThreads), kOrbitThreads, smem, stream>>>(a);
}

}  // namespace lsb
