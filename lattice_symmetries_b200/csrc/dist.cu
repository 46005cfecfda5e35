// dist.cu -- the two hot paths across the GPUs of one node, INSIDE the library: one process per GPU, one NCCL
// communicator per process (created here, not borrowed from the host language), every collective ordered on the
// library stream.
//
// Replaces the multi-locale half of the Chapel driver:
//   hash partition of the states        chapel/src/StatesEnumeration.chpl:198-212, :537-602
//   per-locale blocks of the product    chapel/src/DistributedMatrixVector.chpl:1060-1088 (matrixVectorProduct)
//   radix partition + remote buffers    chapel/src/DistributedMatrixVector.chpl:179-339, :545-579, :775-807
//
// Design (not a port).  Representatives are distributed by CONTIGUOUS SORTED RANGE, not by hash: rank r owns rows
// [bounds[r], bounds[r+1]) of the globally sorted list, so a shard is a sorted array with its own local index, the
// owner of a state is a binary search over world-1 splitters, and x / y are split the same way.
//   build     candidates are dealt out in blocks (small ones first: representatives crowd into the low indices),
//             every rank scans its blocks, ONE all-to-all-v moves each piece to the rank that owns its rows.
//   product   (a) all-gather form, when the replicated structures fit: compact keys (2-4 B per state) + level-1
//             table of the WHOLE basis on every rank -- not the 8-byte representatives -- and the pre-scaled vector
//             n_j x_j replicated by an in-place all-gather-v; then the pull-form kernels run on the local rows.
//             (b) all-to-all form: push records (representative, coefficient) grouped by owner, one grouped
//             ncclSend/ncclRecv exchange per chunk of columns, the owner ranks locally and adds with fp64 atomics.
// Every driver below is written against a Team: the ranks this process drives.  A real run drives one rank and the
// collectives are NCCL calls; the emulated team (tests, one GPU) drives all `world` virtual ranks in lockstep and
// its collectives are device-to-device copies -- the per-rank phases are the same code either way.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: libnccl is loaded at run time, the library does not link it

#include <algorithm>
#include <chrono>
#include <memory>
#include <numeric>
#include <string>

#include "state.hpp"

namespace lsb {

// ---- NCCL, loaded at run time ---------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi &nccl() {
  static NcclApi api;
  if (api.handle != nullptr) return api;
  char const *names[] = {getenv("LS_B200_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (char const *name : names) {
    if (name == nullptr || *name == 0) continue;
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);  // torch's bundled copy when it is already in the process
    if (h != nullptr) break;
  }
  if (h == nullptr) throw CudaFailure{"libnccl.so.2 not found (set LS_B200_NCCL_LIBRARY): multi-GPU paths need NCCL"};
  auto load = [&](auto &fn, char const *symbol) {
    fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, symbol));
    if (fn == nullptr) throw CudaFailure{std::string("libnccl: missing symbol ") + symbol};
  };
  load(api.GetUniqueId, "ncclGetUniqueId");
  load(api.CommInitRank, "ncclCommInitRank");
  load(api.CommDestroy, "ncclCommDestroy");
  load(api.AllGather, "ncclAllGather");
  load(api.AllReduce, "ncclAllReduce");
  load(api.Broadcast, "ncclBroadcast");
  load(api.Send, "ncclSend");
  load(api.Recv, "ncclRecv");
  load(api.GroupStart, "ncclGroupStart");
  load(api.GroupEnd, "ncclGroupEnd");
  load(api.GetErrorString, "ncclGetErrorString");
  api.handle = h;
  return api;
}

static void nccl_check(ncclResult_t r, char const *expr, int line) {
  if (r != ncclSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "NCCL error (%s) at dist.cu:%d: %s", nccl().GetErrorString(r), line, expr);
    throw CudaFailure{buf};
  }
}
#define NCCL_CHECK(expr) ::lsb::nccl_check((expr), #expr, __LINE__)

struct Comm {
  int world = 1, rank = 0;
  ncclComm_t comm = nullptr;
};
static Comm g_comm;

int comm_world() { return g_comm.comm != nullptr ? g_comm.world : 1; }

DistShard::~DistShard() {
  cudaFree(d_splitters);
  block_free(d_xs_full);
  delete global_index;
  delete push;
}

// ---- host-side planning (pure functions; exported for the CPU tests) ---------------------------------------------
constexpr uint64_t kAlign = 32;  // candidate blocks start on a multiple of 32 (one bit-sliced word)

// Blocks of the candidate-index range [0, total): EQUAL blocks of 32 << shift candidates, block b scanned by rank
// b % world.  A representative is the smallest member of its orbit, so representatives -- and the work of finding
// them -- crowd into the low indices (kagome-36: the first 3 % of the range hold 90 % of them); dealing many small
// blocks round-robin gives every rank the same share of every density regime.  A rank scans its whole share as one
// virtual range (build_ranges, CyclicShare), so small blocks cost nothing in launch size.
int dist_block_shift(uint64_t total, int world) {
  uint64_t min_block = uint64_t(1) << 20;
  if (char const *env = getenv("LS_B200_DIST_MIN_BLOCK"))  // tests: small problems cut into many pieces
    min_block = std::max<uint64_t>(kAlign, strtoull(env, nullptr, 10));
  uint64_t const want = std::max<uint64_t>(min_block, total / ((uint64_t)world * 4096));  // <= ~4096 blocks per rank
  int shift = 0;
  while ((uint64_t(32) << (shift + 1)) <= want) ++shift;
  return shift;
}

std::vector<std::pair<uint64_t, uint64_t>> dist_block_plan(uint64_t total, int world) {
  std::vector<std::pair<uint64_t, uint64_t>> plan;
  uint64_t const block = uint64_t(32) << dist_block_shift(total, world);
  for (uint64_t lo = 0; lo < total; lo += block) plan.emplace_back(lo, std::min(total, lo + block));
  return plan;
}

std::vector<int64_t> dist_even_bounds(int64_t dim, int world) {
  std::vector<int64_t> b((size_t)world + 1);
  for (int r = 0; r <= world; ++r) b[(size_t)r] = (int64_t)(((__int128)dim * r) / world);
  return b;
}

struct Piece {
  int64_t begin, length;  // global rows [begin, begin + length) of the sorted list
  int owner;              // the rank that holds it now
};
struct RedistPlan {
  std::vector<int64_t> scount, sdispl, rcount, rdispl;  // [world], in elements
  struct Place {
    int64_t src, dst, length;  // receive-buffer offset -> local row
  };
  std::vector<Place> places;
  int64_t total_send = 0, total_recv = 0;
};

// Pieces (ascending, disjoint, covering [0, dim)) currently held by their owners in piece order; afterwards rank d
// holds rows [bounds[d], bounds[d+1]).  What rank `me` sends (contiguous per destination: its pieces ascend) and where
// what it receives goes.
RedistPlan plan_redistribution(int world, int me, std::vector<Piece> const &pieces, std::vector<int64_t> const &bounds) {
  RedistPlan p;
  p.scount.assign((size_t)world, 0);
  p.sdispl.assign((size_t)world, 0);
  p.rcount.assign((size_t)world, 0);
  p.rdispl.assign((size_t)world, 0);
  int64_t const my_lo = bounds[(size_t)me], my_hi = bounds[(size_t)me + 1];
  for (Piece const &pc : pieces) {
    int64_t const lo = pc.begin, hi = pc.begin + pc.length;
    if (pc.owner == me) {
      for (int d = 0; d < world; ++d) {
        int64_t const a = std::max(lo, bounds[(size_t)d]), b = std::min(hi, bounds[(size_t)d + 1]);
        if (b > a) p.scount[(size_t)d] += b - a;
      }
    }
    int64_t const a = std::max(lo, my_lo), b = std::min(hi, my_hi);
    if (b > a) p.rcount[(size_t)pc.owner] += b - a;
  }
  for (int d = 1; d < world; ++d) {
    p.sdispl[(size_t)d] = p.sdispl[(size_t)d - 1] + p.scount[(size_t)d - 1];
    p.rdispl[(size_t)d] = p.rdispl[(size_t)d - 1] + p.rcount[(size_t)d - 1];
  }
  p.total_send = p.sdispl[(size_t)world - 1] + p.scount[(size_t)world - 1];
  p.total_recv = p.rdispl[(size_t)world - 1] + p.rcount[(size_t)world - 1];
  std::vector<int64_t> running((size_t)world, 0);
  for (Piece const &pc : pieces) {
    int64_t const a = std::max(pc.begin, my_lo), b = std::min(pc.begin + pc.length, my_hi);
    if (b > a) {
      int64_t &run = running[(size_t)pc.owner];
      p.places.push_back({p.rdispl[(size_t)pc.owner] + run, a - my_lo, b - a});
      run += b - a;
    }
  }
  return p;
}

// Contiguous row ranges of (nearly) equal cost: `edges` are ascending row numbers (edges[0] = 0, edges.back() = dim),
// costs[b] the cost of rows [edges[b], edges[b+1]).  Rank r ends at the block boundary closest to (r+1)/world of the
// total (rows of the sorted basis do not cost the same: kagome-36, eight ranks, even split: 0.87x .. 1.12x the mean).
std::vector<int64_t> dist_balanced_bounds(std::vector<int64_t> const &edges, std::vector<double> const &costs, int world,
                                          bool interpolate = false) {
  int64_t const dim = edges.back();
  std::vector<int64_t> bounds((size_t)world + 1, dim);
  bounds[0] = 0;
  double total = 0;
  for (double c : costs) total += c;
  size_t b = 0;
  double acc = 0;
  for (int r = 0; r + 1 < world; ++r) {
    double const target = total * (double)(r + 1) / (double)world;
    while (b < costs.size() && acc + costs[b] <= target) acc += costs[b++];
    int64_t cut = edges[std::min(b, edges.size() - 1)];
    if (interpolate) {
      // inside block b rows cost about the same (they are neighbours in the sorted list): cut it proportionally
      if (b < costs.size() && costs[b] > 0)
        cut = edges[b] + (int64_t)((double)(edges[b + 1] - edges[b]) * ((target - acc) / costs[b]));
    } else if (b < costs.size() && target - acc > acc + costs[b] - target) {
      acc += costs[b++];
      cut = edges[b];
    }
    bounds[(size_t)r + 1] = std::min(dim, std::max(bounds[(size_t)r], cut));
  }
  return bounds;
}

// Row ranges of equal cost in which no rank holds more than `cap` rows.  The cheap rows sit at the low end of the list,
// so it is the first ranks that hit the cap: they are fixed at the cap one by one and the rest of the range is balanced
// over the remaining ranks.
std::vector<int64_t> dist_capped_bounds(std::vector<int64_t> const &edges, std::vector<double> const &piece_costs, int P,
                                        int64_t cap) {
  int64_t const dim = edges.back();
  std::vector<int64_t> fresh = dist_balanced_bounds(edges, piece_costs, P, true);
  for (int r0 = 0; r0 + 1 < P; ++r0) {
    if (fresh[(size_t)r0 + 1] - fresh[(size_t)r0] <= cap) {
      bool ok = true;
      for (int r = r0; r < P; ++r) ok = ok && fresh[(size_t)r + 1] - fresh[(size_t)r] <= cap;
      if (ok) break;
      // (a later rank is over the cap although this one is not: fix this one where it is and look again)
    } else {
      fresh[(size_t)r0 + 1] = fresh[(size_t)r0] + cap;
    }
    // balance rows [fresh[r0 + 1], dim) over ranks r0 + 1 .. P - 1
    int64_t const start = fresh[(size_t)r0 + 1];
    std::vector<int64_t> e{start};
    std::vector<double> c;
    for (size_t b = 0; b + 1 < edges.size(); ++b) {
      if (edges[b + 1] <= start) continue;
      double cost = piece_costs[b];
      if (edges[b] < start) cost *= (double)(edges[b + 1] - start) / (double)(edges[b + 1] - edges[b]);
      e.push_back(edges[b + 1]);
      c.push_back(cost);
    }
    if (c.empty()) {
      for (int r = r0 + 2; r <= P; ++r) fresh[(size_t)r] = dim;
      break;
    }
    for (auto &v : e) v -= start;
    std::vector<int64_t> const rest = dist_balanced_bounds(e, c, P - r0 - 1, true);
    for (int r = r0 + 1; r <= P; ++r) fresh[(size_t)r] = start + rest[(size_t)(r - r0 - 1)];
  }
  // The cap is what a rank's memory holds, so it outranks the balance: after the pass above only the LAST rank can
  // still be over it (costs falling along the list, or all of the cost in a few rows); push its lower boundary up to
  // the cap and let the excess ripple towards rank 0, which cap * P >= dim leaves within the cap as well.
  if ((__int128)cap * P >= dim)
    for (int r = P - 1; r >= 1; --r)
      if (fresh[(size_t)r + 1] - fresh[(size_t)r] > cap) fresh[(size_t)r] = fresh[(size_t)r + 1] - cap;
  return fresh;
}

// ---- the team: the ranks this process drives --------------------------------------------------------------------
__global__ void add_i64_kernel(int64_t *__restrict__ acc, int64_t const *__restrict__ other, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc[i] += other[i];
}
__global__ void add_f64_kernel(double *__restrict__ acc, double const *__restrict__ other, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc[i] += other[i];
}

struct Team {
  int world = 1;
  std::vector<int> members;  // real: {rank}; emulated: 0 .. world-1
  bool emulated = false;

  size_t size() const { return members.size(); }

  // Every rank contributes k values; every rank receives the world x k table (host).  `mine[m]` belongs to members[m].
  std::vector<uint64_t> gather_host(std::vector<std::vector<uint64_t>> const &mine, size_t k) const {
    std::vector<uint64_t> table((size_t)world * k, 0);
    if (emulated) {
      for (size_t m = 0; m < members.size(); ++m)
        std::copy(mine[m].begin(), mine[m].begin() + (ptrdiff_t)k, table.begin() + (ptrdiff_t)((size_t)members[m] * k));
      return table;
    }
    Runtime &rt = runtime();
    static DeviceBuffer<uint64_t> stage;
    uint64_t *d = stage.reserve(table.size());
    CUDA_CHECK(cudaMemcpyAsync(d + (size_t)g_comm.rank * k, mine[0].data(), sizeof(uint64_t) * k, cudaMemcpyHostToDevice,
                               rt.stream));
    NCCL_CHECK(nccl().AllGather(d + (size_t)g_comm.rank * k, d, k, ncclUint64, g_comm.comm, rt.stream));
    CUDA_CHECK(cudaMemcpyAsync(table.data(), d, sizeof(uint64_t) * table.size(), cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    return table;
  }

  // The same for values that live on the device (k per rank).
  std::vector<uint64_t> gather_device(std::vector<unsigned long long const *> const &mine, size_t k) const {
    Runtime &rt = runtime();
    std::vector<uint64_t> table((size_t)world * k, 0);
    if (emulated) {
      for (size_t m = 0; m < members.size(); ++m)
        CUDA_CHECK(cudaMemcpyAsync(table.data() + (size_t)members[m] * k, mine[m], sizeof(uint64_t) * k,
                                   cudaMemcpyDeviceToHost, rt.stream));
      CUDA_CHECK(cudaStreamSynchronize(rt.stream));
      return table;
    }
    static DeviceBuffer<uint64_t> stage;
    uint64_t *d = stage.reserve(table.size());
    NCCL_CHECK(nccl().AllGather(mine[0], d, k, ncclUint64, g_comm.comm, rt.stream));
    CUDA_CHECK(cudaMemcpyAsync(table.data(), d, sizeof(uint64_t) * table.size(), cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    return table;
  }

  // Minimum over all ranks of a host value (emulated: over the members).
  int64_t min_host(std::vector<int64_t> const &mine) const {
    int64_t local = mine[0];
    for (int64_t v : mine) local = std::min(local, v);
    if (emulated) return local;
    Runtime &rt = runtime();
    static DeviceBuffer<int64_t> stage;
    int64_t *d = stage.reserve(1);
    CUDA_CHECK(cudaMemcpyAsync(d, &local, sizeof local, cudaMemcpyHostToDevice, rt.stream));
    NCCL_CHECK(nccl().AllReduce(d, d, 1, ncclInt64, ncclMin, g_comm.comm, rt.stream));
    CUDA_CHECK(cudaMemcpyAsync(&local, d, sizeof local, cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    return local;
  }

  // Personalised exchange; counts and displacements in bytes, [member][peer].
  void all_to_all_v(std::vector<unsigned char const *> const &send, std::vector<std::vector<size_t>> const &sdispl,
                    std::vector<std::vector<size_t>> const &scount, std::vector<unsigned char *> const &recv,
                    std::vector<std::vector<size_t>> const &rdispl, std::vector<std::vector<size_t>> const &rcount) const {
    Runtime &rt = runtime();
    if (emulated) {
      for (size_t dst = 0; dst < members.size(); ++dst)
        for (size_t src = 0; src < members.size(); ++src) {
          size_t const n = scount[src][dst];
          LSB_CHECK(n == rcount[dst][src], "all_to_all_v: send / receive counts disagree");
          if (n > 0)
            CUDA_CHECK(cudaMemcpyAsync(recv[dst] + rdispl[dst][src], send[src] + sdispl[src][dst], n,
                                       cudaMemcpyDeviceToDevice, rt.stream));
        }
      return;
    }
    int const me = g_comm.rank;
    if (scount[0][(size_t)me] > 0)
      CUDA_CHECK(cudaMemcpyAsync(recv[0] + rdispl[0][(size_t)me], send[0] + sdispl[0][(size_t)me], scount[0][(size_t)me],
                                 cudaMemcpyDeviceToDevice, rt.stream));
    NCCL_CHECK(nccl().GroupStart());
    for (int p = 0; p < world; ++p) {
      if (p == me) continue;
      if (scount[0][(size_t)p] > 0)
        NCCL_CHECK(nccl().Send(send[0] + sdispl[0][(size_t)p], scount[0][(size_t)p], ncclChar, p, g_comm.comm, rt.stream));
      if (rcount[0][(size_t)p] > 0)
        NCCL_CHECK(nccl().Recv(recv[0] + rdispl[0][(size_t)p], rcount[0][(size_t)p], ncclChar, p, g_comm.comm, rt.stream));
    }
    NCCL_CHECK(nccl().GroupEnd());
  }

  // In place: segment r = bytes [displs[r], displs[r+1]) of every buffer is valid on rank r; afterwards everywhere.
  void all_gather_v(std::vector<unsigned char *> const &buf, std::vector<size_t> const &displs,
                    cudaStream_t stream = nullptr) const {
    Runtime &rt = runtime();
    if (stream == nullptr) stream = rt.stream;
    if (emulated) {
      for (size_t dst = 0; dst < members.size(); ++dst)
        for (size_t src = 0; src < members.size(); ++src) {
          size_t const r = (size_t)members[src];
          size_t const n = displs[r + 1] - displs[r];
          if (dst != src && n > 0)
            CUDA_CHECK(cudaMemcpyAsync(buf[dst] + displs[r], buf[src] + displs[r], n, cudaMemcpyDeviceToDevice, rt.stream));
        }
      return;
    }
    NCCL_CHECK(nccl().GroupStart());
    for (int root = 0; root < world; ++root) {
      size_t const n = displs[(size_t)root + 1] - displs[(size_t)root];
      if (n > 0)
        NCCL_CHECK(nccl().Broadcast(buf[0] + displs[(size_t)root], buf[0] + displs[(size_t)root], n, ncclChar, root,
                                    g_comm.comm, stream));
    }
    NCCL_CHECK(nccl().GroupEnd());
  }

  void all_reduce_sum_i64(std::vector<int64_t *> const &buf, int64_t n) const {
    Runtime &rt = runtime();
    if (n <= 0) return;
    if (!emulated) {
      NCCL_CHECK(nccl().AllReduce(buf[0], buf[0], (size_t)n, ncclInt64, ncclSum, g_comm.comm, rt.stream));
      return;
    }
    unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
    for (size_t m = 1; m < members.size(); ++m) {
      add_i64_kernel<<<blocks, 256, 0, rt.stream>>>(buf[0], buf[m], n);
      count_launch();
    }
    for (size_t m = 1; m < members.size(); ++m)
      CUDA_CHECK(cudaMemcpyAsync(buf[m], buf[0], sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToDevice, rt.stream));
  }

  void all_reduce_sum_f64(std::vector<double *> const &buf, int64_t n) const {
    Runtime &rt = runtime();
    if (n <= 0) return;
    if (!emulated) {
      NCCL_CHECK(nccl().AllReduce(buf[0], buf[0], (size_t)n, ncclFloat64, ncclSum, g_comm.comm, rt.stream));
      return;
    }
    unsigned const blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8);
    for (size_t m = 1; m < members.size(); ++m) {
      add_f64_kernel<<<blocks, 256, 0, rt.stream>>>(buf[0], buf[m], n);
      count_launch();
    }
    for (size_t m = 1; m < members.size(); ++m)
      CUDA_CHECK(cudaMemcpyAsync(buf[m], buf[0], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, rt.stream));
  }
};

static Team real_team() {
  Team t;
  t.world = comm_world();
  t.members = {g_comm.comm != nullptr ? g_comm.rank : 0};
  t.emulated = g_comm.comm == nullptr;  // a single rank without a communicator: its collectives are no-ops
  return t;
}
static Team emulated_team(int world) {
  Team t;
  t.world = world;
  t.members.resize((size_t)world);
  std::iota(t.members.begin(), t.members.end(), 0);
  t.emulated = true;
  return t;
}

// ---- redistribution of sorted pieces into contiguous row ranges --------------------------------------------------
struct PlaceDev {
  int64_t src, dst, length;
};
// place p: fresh[dst .. dst + length) = recv[src .. src + length)   (8-byte elements; one CTA per place, round-robin)
__global__ void __launch_bounds__(256)
place_kernel(PlaceDev const *__restrict__ places, int64_t n, uint64_t const *__restrict__ recv, uint64_t *__restrict__ fresh) {
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    PlaceDev const pl = places[p];
    for (int64_t i = threadIdx.x; i < pl.length; i += blockDim.x) fresh[pl.dst + i] = recv[pl.src + i];
  }
}

static std::vector<std::unique_ptr<DeviceBuffer<uint64_t>>> &receive_buffers() {
  static std::vector<std::unique_ptr<DeviceBuffer<uint64_t>>> buffers;
  return buffers;
}
// Scratch that grew with the basis is kept for the next build only while it is small.
static void release_redistribute_scratch(size_t above_bytes) {
  for (auto &b : receive_buffers())
    if (b->capacity * sizeof(uint64_t) > above_bytes) b->release();
}

// data[m]: the pieces owned by member m, concatenated (device); replaced by its new rows, in a buffer obtained from
// `allocate(elements)` -- the old buffer is freed.  The receive buffers live for the process (grow-only): with peer
// access enabled every cudaMalloc / cudaFree maps or unmaps on all GPUs and costs milliseconds.
static void redistribute(Team const &team, std::vector<Piece> const &pieces, std::vector<int64_t> const &bounds,
                         std::vector<void *> &data, void *(*allocate)(uint64_t elements)) {
  Runtime &rt = runtime();
  size_t const M = team.size();
  size_t const elem = 8;
  std::vector<RedistPlan> plans;
  std::vector<unsigned char const *> send(M);
  std::vector<unsigned char *> recv(M), fresh(M);
  std::vector<std::vector<size_t>> sd(M), sc(M), rd(M), rc(M);
  auto &receive = receive_buffers();
  while (receive.size() < M) receive.push_back(std::make_unique<DeviceBuffer<uint64_t>>());
  for (size_t m = 0; m < M; ++m) {
    plans.push_back(plan_redistribution(team.world, team.members[m], pieces, bounds));
    RedistPlan const &p = plans.back();
    auto bytes = [&](std::vector<int64_t> const &v) {
      std::vector<size_t> out(v.size());
      for (size_t i = 0; i < v.size(); ++i) out[i] = (size_t)v[i] * elem;
      return out;
    };
    sd[m] = bytes(p.sdispl);
    sc[m] = bytes(p.scount);
    rd[m] = bytes(p.rdispl);
    rc[m] = bytes(p.rcount);
    send[m] = static_cast<unsigned char const *>(data[m]);
    recv[m] = reinterpret_cast<unsigned char *>(receive[m]->reserve((size_t)p.total_recv + 1));
    fresh[m] = static_cast<unsigned char *>(allocate((uint64_t)p.total_recv));
  }
  team.all_to_all_v(send, sd, sc, recv, rd, rc);
  static DeviceBuffer<PlaceDev> table;
  for (size_t m = 0; m < M; ++m) {
    size_t const n = plans[m].places.size();
    if (n == 0) continue;
    if (n <= 8) {
      for (auto const &pl : plans[m].places)
        CUDA_CHECK(cudaMemcpyAsync(fresh[m] + (size_t)pl.dst * elem, recv[m] + (size_t)pl.src * elem,
                                   (size_t)pl.length * elem, cudaMemcpyDeviceToDevice, rt.stream));
      continue;
    }
    static_assert(sizeof(PlaceDev) == sizeof(RedistPlan::Place), "place layout");
    PlaceDev *d = table.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(d, plans[m].places.data(), sizeof(PlaceDev) * n, cudaMemcpyHostToDevice, rt.stream));
    unsigned const blocks = (unsigned)std::min<size_t>(n, (size_t)rt.sm_count * 8);
    place_kernel<<<blocks, 256, 0, rt.stream>>>(d, (int64_t)n, reinterpret_cast<uint64_t const *>(recv[m]),
                                                reinterpret_cast<uint64_t *>(fresh[m]));
    count_launch();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));  // (the table buffer is reused by the next member)
  }
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  for (size_t m = 0; m < M; ++m) {
    block_free(data[m]);
    data[m] = fresh[m];
  }
}

static void release_redistribute_scratch(size_t above_bytes);
static void *allocate_representatives(uint64_t elements) { return alloc_representatives(elements); }
static void *allocate_plain(uint64_t elements) {
  return block_alloc(8 * std::max<uint64_t>(elements, 1), false);
}

// ---- distributed build ----------------------------------------------------------------------------------------------
enum : int { kDistNoGlobalIndex = 1, kDistWideIndex = 2, kDistNoBalance = 4 };

static DistShard *shard_of(ls_hs_basis const *basis) {
  IndexData *ix = index_of(basis);
  return ix != nullptr ? ix->dist : nullptr;
}

// Replicated state -> global row structure.  Small bases (< 2^32 states and a few GB): the representatives themselves
// are all-gathered and the ordinary two-level index is built over them -- identical lookups to the single-GPU path.
// Otherwise ("wide"): only the compact keys are replicated and level 1 is the SUM over ranks of the local
// lower-bound tables (the number of states below a prefix is additive over shards).
static void build_global_index(Team const &team, std::vector<ls_hs_basis *> const &bases, int flags) {
  Runtime &rt = runtime();
  size_t const M = team.size();
  DistShard &first = *shard_of(bases[0]);
  int64_t const dim = first.dim;
  int const number_bits = index_of(bases[0])->number_bits;
  if (dim == 0 || number_bits == 0) return;
  bool wide = (flags & kDistWideIndex) != 0 || dim >= (int64_t(1) << 32) || (size_t)dim * 8 > (size_t(8) << 30);
  if (char const *env = getenv("LS_B200_DIST_INDEX")) wide = strcmp(env, "wide") == 0 ? true : (dim < (int64_t(1) << 32) ? false : wide);
  std::vector<size_t> row_displs((size_t)team.world + 1);
  {
    // Does the all-gather form fit?  It needs the replicated lookup structure and the replicated vector on EVERY rank,
    // next to what the caller still wants for itself (its Krylov vectors: LS_B200_DIST_RESERVE_GB, default four local
    // vectors + 4 GB).  All ranks must come to the same answer: the minimum of the free memory decides.
    int const pb = index_choose_prefix_bits(dim, number_bits);
    size_t const key_bytes_est = number_bits - pb <= 16 ? 2 : 4;
    size_t need = (size_t)dim * 8;  // replicated pre-scaled vector
    need += wide ? (size_t)dim * key_bytes_est + ((size_t(1) << pb) + 1) * 8
                 : (size_t)dim * (8 + key_bytes_est) + ((size_t(1) << pb) + 1) * 20;
    int64_t most_rows = 0;
    for (int r = 0; r < team.world; ++r) most_rows = std::max(most_rows, first.bounds[(size_t)r + 1] - first.bounds[(size_t)r]);
    size_t reserve = (size_t)most_rows * 8 * 4 + (size_t(4) << 30);
    if (char const *env = getenv("LS_B200_DIST_RESERVE_GB")) reserve = (size_t)(atof(env) * (double)(size_t(1) << 30));
    size_t free_bytes = 0, total_bytes = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_bytes, &total_bytes));
    free_bytes += block_cache_idle_bytes();  // (handed back to the driver on demand)
    if (team.emulated) free_bytes /= std::max<size_t>(1, M);  // virtual ranks share one device
    int64_t const headroom = team.min_host({(int64_t)free_bytes - (int64_t)(need + need / 16) - (int64_t)reserve});
    if (headroom < 0) {
      if (getenv("LS_B200_PROFILE") != nullptr)
        fprintf(stderr, "[ls_b200] dist_build: the all-gather form needs %.1f GB per rank (+ %.1f GB reserve), %.1f GB are "
                        "free: products take the all-to-all form\n",
                (double)need / 1e9, (double)reserve / 1e9, (double)free_bytes / 1e9);
      return;
    }
  }
  if (!wide) {
    std::vector<unsigned char *> full(M);
    for (size_t r = 0; r <= (size_t)team.world; ++r) row_displs[r] = (size_t)first.bounds[r] * 8;
    for (size_t m = 0; m < M; ++m) {
      IndexData *local = index_of(bases[m]);
      DistShard &sh = *local->dist;
      void *p = nullptr;
      p = alloc_local((size_t)dim * 8);
      full[m] = static_cast<unsigned char *>(p);
      if (local->number_states > 0)
        CUDA_CHECK(cudaMemcpyAsync(full[m] + row_displs[(size_t)sh.rank], local->d_reps, (size_t)local->number_states * 8,
                                   cudaMemcpyDeviceToDevice, rt.stream));
    }
    team.all_gather_v(full, row_displs);
    for (size_t m = 0; m < M; ++m)
      index_of(bases[m])->dist->global_index =
          create_index_from_device(reinterpret_cast<uint64_t *>(full[m]), dim, number_bits, 22);
    return;
  }
  int prefix_bits = index_choose_prefix_bits(dim, number_bits);
  if (char const *env = getenv("LS_B200_DIST_PREFIX")) prefix_bits = std::max(1, std::min(atoi(env), number_bits));  // A/B knob
  int const shift = number_bits - prefix_bits;
  if (shift > 32) return;  // no compact keys: only the all-to-all form is available
  int const key_bytes = shift <= 16 ? 2 : 4;
  int64_t const number_offsets = (int64_t(1) << prefix_bits) + 1;
  std::vector<unsigned char *> keys(M);
  std::vector<int64_t *> offsets(M);
  for (size_t r = 0; r <= (size_t)team.world; ++r) row_displs[r] = (size_t)first.bounds[r] * (size_t)key_bytes;
  for (size_t m = 0; m < M; ++m) {
    IndexData *local = index_of(bases[m]);
    DistShard &sh = *local->dist;
    void *k = nullptr;
    k = alloc_local((size_t)dim * (size_t)key_bytes);
    keys[m] = static_cast<unsigned char *>(k);
    alloc_local(&offsets[m], sizeof(int64_t) * (size_t)number_offsets);
    index_local_keys(local->d_reps, local->number_states, shift, keys[m] + row_displs[(size_t)sh.rank], key_bytes);
    index_local_offsets64(local->d_reps, local->number_states, shift, number_offsets, offsets[m]);
  }
  team.all_gather_v(keys, row_displs);
  team.all_reduce_sum_i64(offsets, number_offsets);
  for (size_t m = 0; m < M; ++m) {
    auto *g = new IndexData();
    g->number_states = dim;
    g->number_bits = number_bits;
    g->prefix_bits = prefix_bits;
    g->shift = shift;
    g->d_reps = nullptr;
    g->owns_d_reps = false;
    g->d_offsets64 = offsets[m];
    if (key_bytes == 2) g->d_lows16 = reinterpret_cast<uint16_t *>(keys[m]);
    else g->d_lows32 = reinterpret_cast<uint32_t *>(keys[m]);
    g->steps = index_steps_from_offsets64(offsets[m], number_offsets - 1);
    index_of(bases[m])->dist->global_index = g;
  }
}

static void dist_build(Team const &team, std::vector<ls_hs_basis *> const &bases,
                       std::vector<ls_hs_operator const *> const &balance_for, int flags) {
  Runtime &rt = runtime();
  size_t const M = team.size();
  int const P = team.world;
  for (ls_hs_basis *b : bases) LSB_CHECK(b->representatives.num_elts == 0 && index_of(b) == nullptr, "basis is already built");
  static bool const profile = getenv("LS_B200_PROFILE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  std::string phases;
  auto lap = [&](char const *what) {
    if (!profile) return;
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    auto const now = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, " %s %.2f ms,", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    phases += buf;
    t_last = now;
  };
  bool const with_norms = basis_info(bases[0]).has_permutation_symmetries;

  // 1. every rank scans its blocks of the candidate range
  uint64_t const total = number_candidates(bases[0]);
  int const block_shift = dist_block_shift(total, P);
  size_t const nb = (size_t)cyclic_share(total, block_shift, P, 0).blocks_total;
  size_t const per_rank = std::max<size_t>(1, (nb + (size_t)P - 1) / (size_t)P);
  std::vector<std::vector<uint64_t>> counts(M);
  std::vector<void *> reps(M, nullptr), norms(M, nullptr);
  std::vector<uint64_t const *> starts(M, nullptr);
  std::vector<uint64_t *> kept_starts;
  double kernel_ms = 0;
  for (size_t m = 0; m < M; ++m) {
    CyclicShare const share = cyclic_share(total, block_shift, P, team.members[m]);
    BuildResult r = build_ranges(bases[m], Ranges{}, &counts[m], false, &share);
    kernel_ms = std::max(kernel_ms, rt.last_build_ms);
    reps[m] = r.d_reps;
    norms[m] = r.d_norms;
    starts[m] = r.d_block_starts;
    if (M > 1 && r.d_block_starts != nullptr && !counts[m].empty()) {
      // emulated ranks share the library's scratch: keep this member's block starts
      size_t const n = counts[m].size() + 1;
      uint64_t *keep = nullptr;
      CUDA_CHECK(cudaMalloc(&keep, sizeof(uint64_t) * n));
      CUDA_CHECK(cudaMemcpyAsync(keep, r.d_block_starts, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, rt.stream));
      CUDA_CHECK(cudaStreamSynchronize(rt.stream));
      starts[m] = keep;
      kept_starts.push_back(keep);
    }
  }
  lap("scan");
  // 2. everybody learns every block's size -- the pieces in block order are the sorted list -- and, when the rows
  //    are to be balanced for an operator, every block's cost (its matrix elements + 4 row-proportional units per
  //    row: alpha, norm, x, y traffic and the term scan)
  bool const balance = !balance_for.empty() && balance_for[0] != nullptr && (flags & kDistNoBalance) == 0 && P > 1;
  std::vector<std::vector<uint64_t>> costs(M);
  if (balance) {
    static std::vector<std::unique_ptr<DeviceBuffer<uint64_t>>> scratch;
    while (scratch.size() < M) scratch.push_back(std::make_unique<DeviceBuffer<uint64_t>>());
    for (size_t m = 0; m < M; ++m) {
      size_t const mine = counts[m].size();  // blocks of this member (before padding)
      costs[m].assign(per_rank, 0);
      if (mine == 0 || starts[m] == nullptr) continue;
      uint64_t *d = scratch[m]->reserve(mine);
      uint64_t rows_m = 0;
      for (uint64_t c : counts[m]) rows_m += c;
      count_elements_segments(operator_dev(balance_for[m]), static_cast<uint64_t const *>(reps[m]), (int64_t)rows_m, starts[m],
                              (int64_t)mine, d);
      CUDA_CHECK(cudaMemcpyAsync(costs[m].data(), d, sizeof(uint64_t) * mine, cudaMemcpyDeviceToHost, rt.stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    for (size_t m = 0; m < M; ++m)
      for (size_t k = 0; k < counts[m].size(); ++k) costs[m][k] += 4 * counts[m][k];
  }
  for (size_t m = 0; m < M; ++m) counts[m].resize(per_rank, 0);
  std::vector<uint64_t> const table = team.gather_host(counts, per_rank);
  std::vector<uint64_t> cost_table;
  if (balance) cost_table = team.gather_host(costs, per_rank);
  std::vector<Piece> pieces;
  std::vector<int64_t> edges{0};
  std::vector<double> piece_costs;
  int64_t dim = 0;
  for (size_t b = 0; b < nb; ++b) {
    size_t const at = (b % (size_t)P) * per_rank + b / (size_t)P;
    int64_t const n = (int64_t)table[at];
    pieces.push_back({dim, n, (int)(b % (size_t)P)});
    dim += n;
    edges.push_back(dim);
    if (balance) piece_costs.push_back((double)cost_table[at]);
  }
  // 3. row ranges: even, or of equal cost (boundaries interpolated inside a block: neighbouring rows cost the same)
  std::vector<int64_t> bounds = dist_even_bounds(dim, P);
  if (balance && dim >= 4096 * (int64_t)P) bounds = dist_balanced_bounds(edges, piece_costs, P, true);
  for (uint64_t *k : kept_starts) cudaFree(k);
  lap("counts");
  // 4. one all-to-all-v per array: every piece travels to the rank that owns its rows, straight into the final arrays
  redistribute(team, pieces, bounds, reps, allocate_representatives);
  if (with_norms) redistribute(team, pieces, bounds, norms, allocate_plain);
  lap("redistribute");
  // 5. the local rows become the basis' representatives (host view, local index), tagged with the shard layout
  std::vector<std::vector<uint64_t>> firsts(M, std::vector<uint64_t>(1, ~uint64_t(0)));
  for (size_t m = 0; m < M; ++m) {
    int const r = team.members[m];
    int64_t const n = bounds[(size_t)r + 1] - bounds[(size_t)r];
    if (n > 0)
      CUDA_CHECK(cudaMemcpyAsync(firsts[m].data(), reps[m], 8, cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    index_set_lean(getenv("LS_B200_DIST_LEAN_LOCAL") != nullptr);
    install_representatives(bases[m], static_cast<uint64_t *>(reps[m]), with_norms ? static_cast<double *>(norms[m]) : nullptr,
                            (uint64_t)n, 22);
    index_set_lean(false);
    IndexData *local = index_of(bases[m]);
    local->identity = false;  // a shard's rows start at bounds[r]: index == state only holds for the unsharded list
    auto *sh = new DistShard();
    sh->world = P;
    sh->rank = r;
    sh->dim = dim;
    sh->bounds = bounds;
    local->dist = sh;
  }
  std::vector<uint64_t> const splitters = team.gather_host(firsts, 1);
  for (size_t m = 0; m < M; ++m) {
    DistShard &sh = *index_of(bases[m])->dist;
    sh.splitters = splitters;
    // an empty rank owns nothing: give it the splitter of the next non-empty rank so that the owner search skips it
    for (int r = P - 2; r >= 0; --r)
      if (bounds[(size_t)r + 1] == bounds[(size_t)r]) sh.splitters[(size_t)r] = sh.splitters[(size_t)r + 1];
    CUDA_CHECK(cudaMalloc(&sh.d_splitters, sizeof(uint64_t) * (size_t)P));
    CUDA_CHECK(cudaMemcpy(sh.d_splitters, sh.splitters.data(), sizeof(uint64_t) * (size_t)P, cudaMemcpyHostToDevice));
    sh.push = new PushBuffers();
  }
  lap("install + local index");
  release_redistribute_scratch(size_t(1) << 30);
  release_count_scratch(size_t(1) << 30);
  // 6. replicated lookup structure of the all-gather products
  if ((flags & kDistNoGlobalIndex) == 0) build_global_index(team, bases, flags);
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  lap("replicated index");
  if (profile)
    fprintf(stderr, "[ls_b200] dist_build rank %d/%d: %lld states, %zu blocks of %llu candidates;%s\n", team.members[0], P,
            (long long)dim, nb, (unsigned long long)(uint64_t(32) << block_shift), phases.c_str());
  rt.last_build_ms = kernel_ms;
}

// ---- re-balancing by MEASURED cost ------------------------------------------------------------------------------------
// Matrix elements do not all cost the same: the rows at the low end of the sorted list (many leading zeros) couple to
// representatives close to themselves, so their lookups and gathers hit warm cache lines, while the rows at the high
// end scatter over the whole basis.  Kagome-42 on 8 GPUs, rows split by element count: rank + gather takes 2.1 s on
// rank 0 and 5.5 s on rank 7.  Given the measured cost density of one product (costs[m][k] = kernel time of the k-th
// of kCostSegments equal pieces of member m's rows), the row boundaries move so that every rank gets the same TIME.
// Only the local rows move; the replicated index and the layout of the replicated vector refer to GLOBAL rows and
// stay valid.  The caller's vectors must be re-split afterwards (ls_b200_dist_info has the new range).
constexpr int kCostSegments = 1024;

static bool dist_rebalance(Team const &team, std::vector<ls_hs_basis *> const &bases,
                           std::vector<std::vector<double>> const &costs) {
  Runtime &rt = runtime();
  size_t const M = team.size();
  int const P = team.world;
  std::vector<IndexData *> local(M);
  for (size_t m = 0; m < M; ++m) {
    local[m] = index_of(bases[m]);
    LSB_CHECK(local[m] != nullptr && local[m]->dist != nullptr, "the basis was not built by the distributed build");
    LSB_CHECK(costs[m].size() == (size_t)kCostSegments, "dist_rebalance: one cost per segment");
  }
  std::vector<int64_t> const bounds = local[0]->dist->bounds;
  int64_t const dim = local[0]->dist->dim;
  bool const with_norms = local[0]->d_norms != nullptr || basis_info(bases[0]).has_permutation_symmetries;
  // everybody learns everybody's cost density
  std::vector<std::vector<uint64_t>> bits(M, std::vector<uint64_t>((size_t)kCostSegments));
  for (size_t m = 0; m < M; ++m)
    for (int k = 0; k < kCostSegments; ++k) memcpy(&bits[m][(size_t)k], &costs[m][(size_t)k], sizeof(double));
  std::vector<uint64_t> const table = team.gather_host(bits, (size_t)kCostSegments);
  std::vector<int64_t> edges{0};
  std::vector<double> piece_costs;
  for (int r = 0; r < P; ++r) {
    int64_t const n = bounds[(size_t)r + 1] - bounds[(size_t)r];
    for (int k = 0; k < kCostSegments; ++k) {
      int64_t const e = bounds[(size_t)r] + (int64_t)(((__int128)n * (k + 1)) / kCostSegments);
      double c;
      memcpy(&c, &table[(size_t)r * kCostSegments + (size_t)k], sizeof c);
      if (e > edges.back()) {
        edges.push_back(e);
        piece_costs.push_back(c);
      } else if (!piece_costs.empty()) {
        piece_costs.back() += c;
      }
    }
  }
  double total_cost = 0;
  for (double c : piece_costs) total_cost += c;
  if (edges.back() != dim || piece_costs.empty() || !(total_cost > 0)) return false;  // nothing measured: nothing moves
  // No rank may grow beyond what its memory holds next to the replicated structures: LS_B200_DIST_MAX_ROWS, default
  // 12 % more than the largest block so far.  The cheap rows sit at the low end, so it is the first ranks that hit the
  // cap: fix them at the cap one by one and balance the rest of the range over the remaining ranks.
  int64_t cap = 0;
  for (int r = 0; r < P; ++r) cap = std::max(cap, bounds[(size_t)r + 1] - bounds[(size_t)r]);
  cap = cap + cap / 8 + 1;
  if (char const *env = getenv("LS_B200_DIST_MAX_ROWS")) cap = std::max<int64_t>(1, atoll(env));
  std::vector<int64_t> const fresh = dist_capped_bounds(edges, piece_costs, P, cap);
  if (fresh == bounds) return false;
  // detach the local arrays from the bases (the index and the host view go, the arrays travel)
  std::vector<void *> reps(M, nullptr), norms(M, nullptr);
  std::vector<DistShard *> shards(M, nullptr);
  for (size_t m = 0; m < M; ++m) {
    shards[m] = local[m]->dist;
    local[m]->dist = nullptr;
    reps[m] = local[m]->d_reps;
    norms[m] = local[m]->d_norms;
    local[m]->d_norms = nullptr;
    local[m]->owns_d_reps = false;  // (a managed array was to be released by the host view's freer: not called either)
    local[m]->d_reps = nullptr;
    delete local[m];
    bases[m]->kernels->state_index_data = nullptr;
    bases[m]->kernels->state_index_kernel = nullptr;
    bases[m]->representatives.elts = nullptr;
    bases[m]->representatives.num_elts = 0;
    bases[m]->representatives.freer = nullptr;
  }
  std::vector<Piece> held;
  for (int r = 0; r < P; ++r) held.push_back({bounds[(size_t)r], bounds[(size_t)r + 1] - bounds[(size_t)r], r});
  redistribute(team, held, fresh, reps, allocate_representatives);
  if (with_norms) redistribute(team, held, fresh, norms, allocate_plain);
  release_redistribute_scratch(size_t(1) << 30);
  std::vector<std::vector<uint64_t>> firsts(M, std::vector<uint64_t>(1, ~uint64_t(0)));
  for (size_t m = 0; m < M; ++m) {
    int const r = team.members[m];
    int64_t const n = fresh[(size_t)r + 1] - fresh[(size_t)r];
    if (n > 0) CUDA_CHECK(cudaMemcpyAsync(firsts[m].data(), reps[m], 8, cudaMemcpyDeviceToHost, rt.stream));
    CUDA_CHECK(cudaStreamSynchronize(rt.stream));
    index_set_lean(getenv("LS_B200_DIST_LEAN_LOCAL") != nullptr);
    install_representatives(bases[m], static_cast<uint64_t *>(reps[m]), with_norms ? static_cast<double *>(norms[m]) : nullptr,
                            (uint64_t)n, 22);
    index_set_lean(false);
    IndexData *ix = index_of(bases[m]);
    ix->identity = false;
    ix->dist = shards[m];
    shards[m]->bounds = fresh;
  }
  std::vector<uint64_t> const splitters = team.gather_host(firsts, 1);
  for (size_t m = 0; m < M; ++m) {
    DistShard &sh = *shards[m];
    sh.splitters = splitters;
    for (int r = P - 2; r >= 0; --r)
      if (fresh[(size_t)r + 1] == fresh[(size_t)r]) sh.splitters[(size_t)r] = sh.splitters[(size_t)r + 1];
    CUDA_CHECK(cudaMemcpy(sh.d_splitters, sh.splitters.data(), sizeof(uint64_t) * (size_t)P, cudaMemcpyHostToDevice));
  }
  CUDA_CHECK(cudaStreamSynchronize(rt.stream));
  return true;
}

// ---- distributed products ---------------------------------------------------------------------------------------------
enum : int { kProductAuto = 0, kProductAllGather = 1, kProductAllToAll = 2 };

static void dist_matvec(Team const &team, std::vector<ls_hs_operator const *> const &ops,
                        std::vector<double const *> const &x, std::vector<double *> const &y, int mode,
                        bool complex_vectors, cudaEvent_t x_ready = nullptr, double *host_y = nullptr) {
  Runtime &rt = runtime();
  size_t const M = team.size();
  std::vector<IndexData *> local(M);
  for (size_t m = 0; m < M; ++m) {
    local[m] = index_of(ops[m]->basis);
    LSB_CHECK(local[m] != nullptr && local[m]->dist != nullptr, "the basis was not built by the distributed build");
  }
  DistShard const &first = *local[0]->dist;
  if (first.dim == 0) return;
  bool const have_global = first.global_index != nullptr;
  if (mode == kProductAuto) {
    mode = have_global ? kProductAllGather : kProductAllToAll;
    if (char const *env = getenv("LS_B200_DIST_MATVEC")) {
      if (strcmp(env, "alltoall") == 0) mode = kProductAllToAll;
      if (strcmp(env, "allgather") == 0 && have_global) mode = kProductAllGather;
    }
  }
  if (mode == kProductAllGather) {
    LSB_CHECK(have_global, "all-gather products need the replicated index (built without it)");
    size_t const scalar = complex_vectors ? 2 : 1;
    std::vector<unsigned char *> full(M);
    std::vector<size_t> displs((size_t)team.world + 1);
    for (size_t r = 0; r <= (size_t)team.world; ++r) displs[r] = (size_t)first.bounds[r] * 8 * scalar;
    for (size_t m = 0; m < M; ++m) {
      DistShard &sh = *local[m]->dist;
      size_t const words = (size_t)sh.dim * scalar;
      if (sh.xs_full_words < words) {
        block_free(sh.d_xs_full);
        sh.d_xs_full = nullptr;
        alloc_local(&sh.d_xs_full, sizeof(double) * words);
        sh.xs_full_words = words;
      }
      full[m] = reinterpret_cast<unsigned char *>(sh.d_xs_full);
    }
    // Real ranks: pre-scaling (n_j x_j of the local rows, straight into their slot of the replicated vector) and the
    // all-gather run on their own stream, under the canonicalisation of the first chunk of rows (which reads only the
    // representatives); the first kernel that reads the replicated vector waits for the collective.
    cudaEvent_t gathered = nullptr;
    cudaStream_t side = rt.stream;
    bool const overlap = !team.emulated && team.world > 1 && getenv("LS_B200_DIST_NO_OVERLAP") == nullptr;
    if (overlap) {
      static cudaStream_t comm_stream = nullptr;
      static cudaEvent_t inputs_ready = nullptr, done = nullptr;
      if (comm_stream == nullptr) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&inputs_ready, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      }
      CUDA_CHECK(cudaEventRecord(inputs_ready, rt.stream));  // x written by earlier work on the library stream; the
      CUDA_CHECK(cudaStreamWaitEvent(comm_stream, inputs_ready, 0));  // previous product is through with the replica
      side = comm_stream;
      gathered = done;
    }
    if (x_ready != nullptr) CUDA_CHECK(cudaStreamWaitEvent(side, x_ready, 0));
    // LS_B200_PROFILE: device time of pre-scaling + all-gather, reported by the NEXT product (no extra sync)
    static bool const profile = getenv("LS_B200_PROFILE") != nullptr;
    static cudaEvent_t t_begin = nullptr, t_end = nullptr;
    static bool timed = false;
    if (profile && !team.emulated) {
      if (t_begin == nullptr) {
        CUDA_CHECK(cudaEventCreate(&t_begin));
        CUDA_CHECK(cudaEventCreate(&t_end));
      } else if (timed && cudaEventQuery(t_end) == cudaSuccess) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t_begin, t_end) == cudaSuccess) runtime().last_allgather_ms = ms;
      }
      CUDA_CHECK(cudaEventRecord(t_begin, side));
    }
    for (size_t m = 0; m < M; ++m) {
      DistShard &sh = *local[m]->dist;
      launch_prescale(local[m]->number_states, complex_vectors, local[m]->d_norms, x[m],
                      sh.d_xs_full + (size_t)sh.bounds[(size_t)sh.rank] * scalar, side);
    }
    team.all_gather_v(full, displs, side);
    if (profile && !team.emulated) {
      CUDA_CHECK(cudaEventRecord(t_end, side));
      timed = true;
    }
    if (overlap) CUDA_CHECK(cudaEventRecord(gathered, side));
    for (size_t m = 0; m < M; ++m) {
      DistShard &sh = *local[m]->dist;
      MvTarget t{};
      t.xs_ready = gathered;
      t.index = sh.global_index->view();
      t.rows = local[m]->d_reps;
      t.norms = local[m]->d_norms;
      t.number_rows = local[m]->number_states;
      t.xs = sh.d_xs_full;
      if (t.number_rows > 0)
        matvec_device(ops[m], 0, t.number_rows, x[m], y[m], complex_vectors, 1, 0, 0, host_y, 0, &t);
    }
    // (a rank without rows launched nothing that waits: later work on the library stream must still not touch the
    // replicated vector before the collective is through)
    if (gathered != nullptr) CUDA_CHECK(cudaStreamWaitEvent(rt.stream, gathered, 0));
    return;
  }
  LSB_CHECK(!complex_vectors, "all-to-all products take real vectors (as the reference's matvec)");
  if (x_ready != nullptr) CUDA_CHECK(cudaStreamWaitEvent(rt.stream, x_ready, 0));
  LSB_CHECK(team.world <= 64, "at most 64 ranks");
  int64_t const chunk_rows = push_chunk_rows(ops[0], 0);
  int64_t rounds = 0;
  for (int r = 0; r < team.world; ++r)
    rounds = std::max(rounds, (first.bounds[(size_t)r + 1] - first.bounds[(size_t)r] + chunk_rows - 1) / chunk_rows);
  for (size_t m = 0; m < M; ++m) push_begin(ops[m], *local[m], x[m], y[m]);
  size_t const P = (size_t)team.world;
  for (int64_t c = 0; c < rounds; ++c) {
    std::vector<unsigned long long const *> displs(M);
    for (size_t m = 0; m < M; ++m) {
      int64_t const begin = c * chunk_rows;
      int64_t const nrows = std::max<int64_t>(0, std::min(chunk_rows, local[m]->number_states - begin));
      push_produce(ops[m], *local[m], *local[m]->dist, begin, nrows, x[m]);
      displs[m] = local[m]->dist->push->displs.ptr;
    }
    std::vector<uint64_t> const table = team.gather_device(displs, P + 1);  // [rank][owner] record displacements
    std::vector<unsigned char const *> send(M);
    std::vector<unsigned char *> recv(M);
    std::vector<std::vector<size_t>> sd(M, std::vector<size_t>(P)), sc(M, std::vector<size_t>(P)),
        rd(M, std::vector<size_t>(P)), rc(M, std::vector<size_t>(P));
    std::vector<size_t> received(M, 0);
    for (size_t m = 0; m < M; ++m) {
      size_t const r = (size_t)team.members[m];
      size_t at = 0;
      for (size_t p = 0; p < P; ++p) {
        sd[m][p] = (size_t)table[r * (P + 1) + p] * sizeof(PushRecord);
        sc[m][p] = (size_t)(table[r * (P + 1) + p + 1] - table[r * (P + 1) + p]) * sizeof(PushRecord);
        rc[m][p] = (size_t)(table[p * (P + 1) + r + 1] - table[p * (P + 1) + r]) * sizeof(PushRecord);
        rd[m][p] = at;
        at += rc[m][p];
      }
      received[m] = at / sizeof(PushRecord);
      PushBuffers &pb = *local[m]->dist->push;
      send[m] = reinterpret_cast<unsigned char const *>(pb.send.ptr);
      recv[m] = reinterpret_cast<unsigned char *>(pb.recv.reserve(received[m] + 1));
    }
    team.all_to_all_v(send, sd, sc, recv, rd, rc);
    for (size_t m = 0; m < M; ++m)
      push_consume(ops[m], *local[m], local[m]->dist->push->recv.ptr, (int64_t)received[m], y[m]);
  }
  if (host_y != nullptr && local[0]->number_states > 0)
    CUDA_CHECK(cudaMemcpyAsync(host_y, y[0], sizeof(double) * (size_t)local[0]->number_states, cudaMemcpyDeviceToHost, rt.stream));
}

// Entry points for the rest of the library: the product of a basis that carries a shard layout.
bool dist_is_sharded(ls_hs_basis const *basis) { return shard_of(basis) != nullptr; }

void dist_matvec_local(ls_hs_operator const *op, double const *d_x, double *d_y, int mode, bool complex_vectors,
                       cudaEvent_t x_ready, double *host_y) {
  DistShard *sh = shard_of(op->basis);
  LSB_CHECK(sh != nullptr, "not a distributed basis");
  LSB_CHECK(sh->world == comm_world(), "the basis was sharded over a different communicator");
  Team const team = real_team();
  dist_matvec(team, {op}, {d_x}, {d_y}, mode, complex_vectors, x_ready, host_y);
}

void dist_build_local(ls_hs_basis *basis, ls_hs_operator const *balance_for, int flags) {
  Team const team = real_team();
  dist_build(team, {basis}, {balance_for}, flags);
}

}  // namespace lsb

using namespace lsb;

extern "C" {

int ls_b200_comm_unique_id(void *out, size_t bytes) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(out != nullptr && bytes >= sizeof(ncclUniqueId), "ls_b200_comm_unique_id: need a 128-byte buffer");
    ncclUniqueId id;
    NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(out, &id, sizeof id);
    status = 0;
  });
  return status;
}

int ls_b200_comm_init(int world, int rank, void const *unique_id) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(world >= 1 && rank >= 0 && rank < world, "ls_b200_comm_init: invalid world / rank");
    if (g_comm.comm != nullptr) {
      LSB_CHECK(g_comm.world == world && g_comm.rank == rank, "a different communicator is already active");
      status = 0;
      return;
    }
    if (world == 1) {  // nothing to talk to
      status = 0;
      return;
    }
    LSB_CHECK(unique_id != nullptr, "ls_b200_comm_init: unique id missing");
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm = nullptr;
    NCCL_CHECK(nccl().CommInitRank(&comm, world, id, rank));
    g_comm.comm = comm;
    g_comm.world = world;
    g_comm.rank = rank;
    status = 0;
  });
  return status;
}

void ls_b200_comm_finalize(void) {
  if (g_comm.comm == nullptr) return;
  std::lock_guard<std::mutex> lock(runtime().mutex);
  cudaStreamSynchronize(runtime().stream);
  nccl().CommDestroy(g_comm.comm);
  g_comm = Comm{};
}

int ls_b200_comm_size(void) { return comm_world(); }
int ls_b200_comm_rank(void) { return g_comm.comm != nullptr ? g_comm.rank : 0; }

int ls_b200_comm_allreduce_f64(double *values_dev, int count) {
  int status = -1;
  guarded(__func__, [&] {
    if (g_comm.comm != nullptr && count > 0)
      NCCL_CHECK(nccl().AllReduce(values_dev, values_dev, (size_t)count, ncclFloat64, ncclSum, g_comm.comm, runtime().stream));
    status = 0;
  });
  return status;
}

int ls_b200_dist_build(ls_hs_basis *basis, ls_hs_operator const *balance_for, int flags) {
  int status = -1;
  guarded(__func__, [&] {
    dist_build_local(basis, balance_for, flags);
    status = 0;
  });
  return status;
}

int ls_b200_dist_matvec(ls_hs_operator const *op, double const *x_local_dev, double *y_local_dev, int mode) {
  int status = -1;
  guarded(__func__, [&] {
    dist_matvec_local(op, x_local_dev, y_local_dev, mode, false);
    status = 0;
  });
  return status;
}

int ls_b200_dist_matvec_c128(ls_hs_operator const *op, ls_hs_scalar const *x_local_dev, ls_hs_scalar *y_local_dev,
                             int mode) {
  int status = -1;
  guarded(__func__, [&] {
    dist_matvec_local(op, reinterpret_cast<double const *>(x_local_dev), reinterpret_cast<double *>(y_local_dev), mode, true);
    status = 0;
  });
  return status;
}

int ls_b200_dist_rebalance(ls_hs_basis *basis) {
  int status = -1;
  guarded(__func__, [&] {
    IndexData *ix = index_of(basis);
    LSB_CHECK(ix != nullptr && ix->dist != nullptr, "ls_b200_dist_rebalance: not a distributed basis");
    std::vector<std::vector<double>> costs(1, std::vector<double>((size_t)kCostSegments, 0.0));
    // a rank whose last product recorded no costs (no rows, or LS_B200_PROFILE unset) reports zeros: harmless when
    // every rank does so (nothing moves), but all ranks must make this call together either way
    (void)matvec_last_row_costs(ix->number_states, kCostSegments, costs[0].data());
    status = dist_rebalance(real_team(), {basis}, costs) ? 0 : 1;
  });
  return status;
}

int ls_b200_emu_rebalance(ls_hs_basis **bases, int world, double const *costs) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(world >= 1 && bases != nullptr && costs != nullptr, "ls_b200_emu_rebalance: invalid arguments");
    std::vector<ls_hs_basis *> b(bases, bases + world);
    std::vector<std::vector<double>> c((size_t)world);
    for (int r = 0; r < world; ++r) c[(size_t)r].assign(costs + (size_t)r * kCostSegments, costs + (size_t)(r + 1) * kCostSegments);
    status = dist_rebalance(emulated_team(world), b, c) ? 0 : 1;
  });
  return status;
}

int ls_b200_dist_info(ls_hs_basis const *basis, int64_t out[8]) {
  DistShard const *sh = shard_of(basis);
  if (sh == nullptr || out == nullptr) return -1;
  out[0] = sh->world;
  out[1] = sh->rank;
  out[2] = sh->dim;
  out[3] = sh->bounds[(size_t)sh->rank];
  out[4] = sh->bounds[(size_t)sh->rank + 1];
  out[5] = sh->global_index == nullptr ? 0 : (sh->global_index->d_offsets64 != nullptr ? 2 : 1);
  out[6] = sh->global_index != nullptr ? sh->global_index->steps : 0;
  out[7] = sh->global_index != nullptr ? sh->global_index->prefix_bits : 0;
  return 0;
}

int ls_b200_dist_bounds(ls_hs_basis const *basis, int64_t *bounds, int capacity) {
  DistShard const *sh = shard_of(basis);
  if (sh == nullptr || bounds == nullptr || capacity < sh->world + 1) return -1;
  for (int r = 0; r <= sh->world; ++r) bounds[r] = sh->bounds[(size_t)r];
  return sh->world;
}

// ---- emulated team: `world` virtual ranks on the current device, driven in lockstep (tests; also a way to run a
// basis through the sharded code paths on one GPU) --------------------------------------------------------------------
int ls_b200_emu_build(ls_hs_basis **bases, ls_hs_operator const *const *balance_for, int world, int flags) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(world >= 1 && bases != nullptr, "ls_b200_emu_build: invalid arguments");
    std::vector<ls_hs_basis *> b(bases, bases + world);
    std::vector<ls_hs_operator const *> ops;
    if (balance_for != nullptr) ops.assign(balance_for, balance_for + world);
    dist_build(emulated_team(world), b, ops, flags);
    status = 0;
  });
  return status;
}

int ls_b200_emu_matvec(ls_hs_operator const *const *ops, int world, double const *const *x_dev, double *const *y_dev,
                       int mode, int complex_vectors) {
  int status = -1;
  guarded(__func__, [&] {
    LSB_CHECK(world >= 1 && ops != nullptr, "ls_b200_emu_matvec: invalid arguments");
    std::vector<ls_hs_operator const *> o(ops, ops + world);
    std::vector<double const *> x(x_dev, x_dev + world);
    std::vector<double *> y(y_dev, y_dev + world);
    dist_matvec(emulated_team(world), o, x, y, mode, complex_vectors != 0);
    status = 0;
  });
  return status;
}

// ---- planning functions, exported for the CPU tests (no device needed) --------------------------------------------
int64_t ls_b200_plan_blocks(uint64_t total, int world, uint64_t *begins, uint64_t *ends, int64_t capacity) {
  auto const plan = dist_block_plan(total, world);
  if (begins != nullptr && ends != nullptr)
    for (size_t b = 0; b < plan.size() && (int64_t)b < capacity; ++b) {
      begins[b] = plan[b].first;
      ends[b] = plan[b].second;
    }
  return (int64_t)plan.size();
}

// pieces: lengths[i] rows held by owners[i], ascending global order.  Outputs: send / receive counts and displacements
// [world] (elements), places[3 k] = (receive offset, local row, length); returns the number of places (or -1 when
// `capacity` places do not suffice).
int64_t ls_b200_plan_redistribution(int world, int me, int64_t number_pieces, int64_t const *lengths, int32_t const *owners,
                                    int64_t const *bounds, int64_t *scount, int64_t *sdispl, int64_t *rcount,
                                    int64_t *rdispl, int64_t *places, int64_t capacity) {
  std::vector<Piece> pieces;
  int64_t at = 0;
  for (int64_t i = 0; i < number_pieces; ++i) {
    pieces.push_back({at, lengths[i], (int)owners[i]});
    at += lengths[i];
  }
  std::vector<int64_t> b(bounds, bounds + world + 1);
  RedistPlan const p = plan_redistribution(world, me, pieces, b);
  for (int d = 0; d < world; ++d) {
    scount[d] = p.scount[(size_t)d];
    sdispl[d] = p.sdispl[(size_t)d];
    rcount[d] = p.rcount[(size_t)d];
    rdispl[d] = p.rdispl[(size_t)d];
  }
  if ((int64_t)p.places.size() > capacity) return -1;
  for (size_t k = 0; k < p.places.size(); ++k) {
    places[3 * k] = p.places[k].src;
    places[3 * k + 1] = p.places[k].dst;
    places[3 * k + 2] = p.places[k].length;
  }
  return (int64_t)p.places.size();
}

// Row ranges of equal cost with at most `cap` rows per rank (what ls_b200_dist_rebalance computes from measured costs);
// boundaries are interpolated inside a block.
int ls_b200_plan_rebalance_bounds(int64_t number_blocks, int64_t const *edges, double const *costs, int world, int64_t cap,
                                  int64_t *bounds) {
  std::vector<int64_t> e(edges, edges + number_blocks + 1);
  std::vector<double> c(costs, costs + number_blocks);
  auto const b = dist_capped_bounds(e, c, world, cap);
  for (int r = 0; r <= world; ++r) bounds[r] = b[(size_t)r];
  return 0;
}

// A rank's block-cyclic share of the candidate range as the sharded build scans it: out = {block shift (blocks hold
// 32 << shift candidates), blocks over all ranks, blocks of this rank, candidates of this rank}.
int ls_b200_plan_cyclic_share(uint64_t total, int world, int rank, uint64_t out[4]) {
  int const shift = dist_block_shift(total, world);
  CyclicShare const sh = cyclic_share(total, shift, world, rank);
  out[0] = (uint64_t)shift;
  out[1] = sh.blocks_total;
  out[2] = sh.number_blocks;
  out[3] = sh.virtual_candidates;
  return 0;
}

int ls_b200_plan_balanced_bounds(int64_t number_blocks, int64_t const *edges, double const *costs, int world,
                                 int64_t *bounds) {
  std::vector<int64_t> e(edges, edges + number_blocks + 1);
  std::vector<double> c(costs, costs + number_blocks);
  auto const b = dist_balanced_bounds(e, c, world);
  for (int r = 0; r <= world; ++r) bounds[r] = b[(size_t)r];
  return 0;
}

}  // extern "C"
