"""Several GPUs of one node: one process per GPU, the sharding itself lives INSIDE the C library
(csrc/dist.cu) behind the reference's own entry points.

The reference distributes basis states over Chapel locales by *hash*
(chapel/src/StatesEnumeration.chpl:198-212) and pushes (state, coefficient) pairs to their owners through
GASNet PUTs (chapel/src/DistributedMatrixVector.chpl:226-339, :545-579); each locale then works on its own
block of the representatives and of x / y (:1060-1088).  Here a rank owns a *contiguous range of the sorted
representatives*; once ``init_communicator`` has run,

* ``basis.build()`` (= ``ls_hs_build_representatives``) scans this rank's blocks of the candidate range and
  one all-to-all-v leaves every rank with its range -- ``basis.states`` is the local block;
* ``operator.apply_to_state_vector`` (= ``ls_chpl_matrix_vector_product``) takes the local blocks of x and y;
* :class:`DistributedOperator` is the same product with device-resident local vectors, in either form:
  *all-gather* (compact keys of the whole basis + the pre-scaled vector replicated, pull-form kernels) or
  *all-to-all* ((representative, coefficient) records grouped by owner, mirroring the Chapel design).

``torch.distributed`` is only plumbing here: it carries the 128-byte NCCL id to the other ranks; the
collectives on the data path are issued by the library on its own communicator and stream.

:class:`EmulatedRanks` drives the same code for ``world`` *virtual* ranks on one GPU (device-to-device
copies in place of NCCL) -- that is how the single-GPU test tier covers the multi-rank algorithms.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "shard_bounds", "row_bounds", "plan_blocks", "plan_redistribution", "even_bounds", "balanced_bounds",
    "rebalance_bounds", "cyclic_share",
    "init_process", "init_communicator", "Layout", "build_distributed", "rebalance_distributed", "DistributedOperator", "EmulatedRanks",
    "hashed_vector", "hashed_values", "layout_of", "ALLGATHER", "ALLTOALL", "AUTO", "NO_GLOBAL_INDEX", "WIDE_INDEX", "NO_BALANCE",
]

ALIGN = 32  # candidate shards start on a multiple of 32 (one bit-sliced word)
AUTO, ALLGATHER, ALLTOALL = 0, 1, 2                    # product forms (ls_b200_dist_matvec ``mode``)
NO_GLOBAL_INDEX, WIDE_INDEX, NO_BALANCE = 1, 2, 4      # build flags (ls_b200_dist_build ``flags``)


# ---- pure planning (host) ----------------------------------------------------------------------------
def shard_bounds(total: int, world: int, rank: int, align: int = ALIGN) -> Tuple[int, int]:
    """Contiguous candidate-index range of ``rank`` for ``ls_b200_build_shard``: boundaries are multiples of
    ``align``, sizes differ by at most ``align``, the union is [0, total)."""
    words = (total + align - 1) // align
    lo = (words * rank) // world * align
    hi = (words * (rank + 1)) // world * align
    return min(lo, total), min(hi, total)


def row_bounds(dim: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Even row split with a fixed chunk = ceil(dim / world): (row_begin, row_end, chunk)."""
    chunk = (dim + world - 1) // world if dim > 0 else 0
    lo = min(rank * chunk, dim)
    hi = min((rank + 1) * chunk, dim)
    return lo, hi, chunk


def even_bounds(dim: int, world: int) -> List[int]:
    """Row boundaries of the sharded basis before balancing: rank r owns [b[r], b[r+1])."""
    return [(dim * r) // world for r in range(world + 1)]


def plan_blocks(total: int, world: int) -> List[Tuple[int, int]]:
    """Blocks of the candidate-index range, block b scanned by rank b % world (csrc/dist.cu ``dist_block_plan``):
    equal blocks of a power of two >= 2^20 candidates, at most ~4096 per rank -- representatives crowd into the low
    indices, dealing many small blocks round-robin gives every rank the same share of every density regime."""
    from ._lib import lib
    n = int(lib.ls_b200_plan_blocks(int(total), int(world), None, None, 0))
    begins = (C.c_uint64 * max(n, 1))()
    ends = (C.c_uint64 * max(n, 1))()
    lib.ls_b200_plan_blocks(int(total), int(world), begins, ends, n)
    return [(int(begins[i]), int(ends[i])) for i in range(n)]


@dataclass
class RedistributionPlan:
    scount: np.ndarray
    sdispl: np.ndarray
    rcount: np.ndarray
    rdispl: np.ndarray
    places: np.ndarray  # [k, 3]: (offset in the receive buffer, local row, length)


def plan_redistribution(world: int, me: int, lengths: Sequence[int], owners: Sequence[int],
                        bounds: Sequence[int]) -> RedistributionPlan:
    """All-to-all-v plan of rank ``me`` that turns sorted pieces (``lengths[i]`` rows held by ``owners[i]``, in
    ascending order) into contiguous row ranges ``bounds`` (csrc/dist.cu ``plan_redistribution``)."""
    from ._lib import lib, i64_p
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    owners = np.ascontiguousarray(owners, dtype=np.int32)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    out = [np.zeros(world, dtype=np.int64) for _ in range(4)]
    capacity = len(lengths) + world + 1
    places = np.zeros((capacity, 3), dtype=np.int64)
    n = int(lib.ls_b200_plan_redistribution(
        world, me, len(lengths), lengths.ctypes.data_as(i64_p), owners.ctypes.data_as(C.POINTER(C.c_int32)),
        bounds.ctypes.data_as(i64_p), *[o.ctypes.data_as(i64_p) for o in out], places.ctypes.data_as(i64_p), capacity))
    if n < 0:
        raise RuntimeError("ls_b200_plan_redistribution: place table too small")
    return RedistributionPlan(*out, places[:n].copy())


def balanced_bounds(edges: Sequence[int], costs: Sequence[float], world: int) -> List[int]:
    """Contiguous row ranges of (nearly) equal cost; ``costs[b]`` is the cost of rows [edges[b], edges[b+1])."""
    from ._lib import lib, i64_p, f64_p
    edges = np.ascontiguousarray(edges, dtype=np.int64)
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    assert len(edges) == len(costs) + 1
    out = np.zeros(world + 1, dtype=np.int64)
    lib.ls_b200_plan_balanced_bounds(len(costs), edges.ctypes.data_as(i64_p), costs.ctypes.data_as(f64_p), world,
                                     out.ctypes.data_as(i64_p))
    return [int(b) for b in out]


def rebalance_bounds(edges: Sequence[int], costs: Sequence[float], world: int, cap: int) -> List[int]:
    """Row ranges of equal (measured) cost with at most ``cap`` rows per rank, boundaries interpolated inside a block:
    what ``ls_b200_dist_rebalance`` computes (csrc/dist.cu ``dist_capped_bounds``)."""
    from ._lib import lib, i64_p, f64_p
    edges = np.ascontiguousarray(edges, dtype=np.int64)
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    assert len(edges) == len(costs) + 1
    out = np.zeros(world + 1, dtype=np.int64)
    lib.ls_b200_plan_rebalance_bounds(len(costs), edges.ctypes.data_as(i64_p), costs.ctypes.data_as(f64_p), world, int(cap),
                                      out.ctypes.data_as(i64_p))
    return [int(b) for b in out]


def cyclic_share(total: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """(block shift, blocks over all ranks, blocks of ``rank``, candidates of ``rank``) of the sharded build's
    block-cyclic scan (csrc/basis_build.cu ``cyclic_share``): block b -- 32 << shift candidates -- belongs to rank b % world."""
    from ._lib import lib, u64_p
    out = (C.c_uint64 * 4)()
    lib.ls_b200_plan_cyclic_share(int(total), int(world), int(rank), out)
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


# ---- process set-up ------------------------------------------------------------------------------------
_stream = None


def init_process(local_rank: Optional[int] = None):
    """Select this rank's GPU and order all library work on a dedicated torch stream, so that torch's vector
    algebra and the library's kernels / collectives interleave without host syncs."""
    import os
    import torch
    from . import _lib
    global _stream
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.ensure_init()
    if _stream is None:
        _stream = torch.cuda.Stream()
        _lib.lib.ls_b200_set_stream(_stream.cuda_stream)
        _lib.check_error()
    torch.cuda.set_stream(_stream)
    return _stream


def init_communicator(group=None) -> Tuple[int, int]:
    """Create the library's own NCCL communicator over the ranks of the (already initialised)
    ``torch.distributed`` group: rank 0 draws the id, ``torch.distributed`` carries its 128 bytes -- nothing
    else.  Returns (world, rank).  A single process needs no communicator."""
    import torch.distributed as dist
    from . import _lib
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return 1, 0
    if _lib.lib.ls_b200_comm_size() == world:
        return world, rank
    blob = [None]
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        if _lib.lib.ls_b200_comm_unique_id(buf, 128) != 0:
            _lib.check_error()
            raise RuntimeError("ls_b200_comm_unique_id failed")
        blob[0] = bytes(buf)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(blob, src=src, group=group)
    raw = (C.c_ubyte * 128).from_buffer_copy(blob[0])
    status = _lib.lib.ls_b200_comm_init(world, rank, raw)
    _lib.check_error()
    if status != 0:
        raise RuntimeError("ls_b200_comm_init failed")
    return world, rank


# ---- the sharded basis ------------------------------------------------------------------------------------
@dataclass
class Layout:
    world: int
    rank: int
    dim: int            # representatives over all ranks
    row_begin: int      # this rank owns rows [row_begin, row_end) of the globally sorted list
    row_end: int
    global_index: int   # replicated index of the all-gather form: 0 none, 1 two-level, 2 wide (compact keys only)
    search_steps: int
    prefix_bits: int
    bounds: List[int]

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin


def layout_of(basis) -> Layout:
    from ._lib import lib
    out = (C.c_int64 * 8)()
    if lib.ls_b200_dist_info(C.byref(basis._payload), out) != 0:
        raise ValueError("the basis is not sharded (build it with build_distributed, or under init_communicator)")
    world = int(out[0])
    b = (C.c_int64 * (world + 1))()
    lib.ls_b200_dist_bounds(C.byref(basis._payload), b, world + 1)
    return Layout(world, int(out[1]), int(out[2]), int(out[3]), int(out[4]), int(out[5]), int(out[6]), int(out[7]),
                  [int(v) for v in b])


def build_distributed(basis, balance_for=None, flags: int = 0) -> Layout:
    """``ls_b200_dist_build``: the sharded build with its options spelled out (``basis.build()`` under an active
    communicator is the same with the defaults).  ``balance_for``: an Operator on this basis; rows are then
    split by its matrix-element count instead of evenly."""
    from . import _lib
    op = C.byref(balance_for._payload) if balance_for is not None else None
    status = _lib.lib.ls_b200_dist_build(C.byref(basis._payload), op, int(flags))
    _lib.check_error()
    if status != 0:
        raise RuntimeError("ls_b200_dist_build failed")
    return layout_of(basis)


def rebalance_distributed(basis) -> Layout:
    """``ls_b200_dist_rebalance``: move the row boundaries so that every rank spends the same time on a product,
    from the kernel times measured during the last product (``LS_B200_PROFILE=1`` must be set before the first
    product).  Collective; vectors split by the old layout must be re-split.  Returns the new layout."""
    from . import _lib
    status = _lib.lib.ls_b200_dist_rebalance(C.byref(basis._payload))
    _lib.check_error()
    if status < 0:
        raise RuntimeError("ls_b200_dist_rebalance failed")
    basis._keep.pop("views", None)
    return layout_of(basis)


def _as_int64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _splitmix_unit(z):
    """splitmix64 finaliser of int64 tensor ``z`` (wrapping arithmetic) -> float64 in [0, 1)."""
    import torch
    z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * -4658895280553007687      # 0xBF58476D1CE4E5B9
    z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * -7723592293110705685     # 0x94D049BB133111EB
    z = z ^ ((z >> 31) & 0x1FFFFFFFF)
    return ((z >> 11) & 0x1FFFFFFFFFFFFF).to(torch.float64) * (1.0 / (1 << 53))


def hashed_values(rows, seed: int = 42):
    """Entries of the hashed vector at arbitrary GLOBAL rows (an int64 tensor, any device)."""
    import torch
    z = rows.to(torch.int64) + _as_int64((seed + 1) * 0x9E3779B97F4A7C15)   # golden-ratio increment
    return 2.0 * _splitmix_unit(z) - 1.0


def hashed_vector(row_begin: int, row_end: int, seed: int = 42, device="cuda"):
    """Deterministic pseudo-random vector entries in [-1, 1) that depend only on (seed, GLOBAL row): every rank
    fills its own rows, and the vector is the same for any number of ranks.  (A start vector for Lanczos on a
    basis whose full length -- 9.6e9 for kagome-42 -- no single host could draw.)"""
    import torch
    out = torch.empty(row_end - row_begin, dtype=torch.float64, device=device)
    chunk = 1 << 25
    for lo in range(row_begin, row_end, chunk):
        hi = min(row_end, lo + chunk)
        out[lo - row_begin:hi - row_begin] = hashed_values(torch.arange(lo, hi, dtype=torch.int64, device=device), seed)
    return out


class DistributedOperator:
    """y_local = (H x)_local on a sharded basis, vectors device-resident (torch tensors of local length)."""

    device = "cuda"

    def __init__(self, operator, mode: int = AUTO):
        self.op = operator
        self.mode = mode
        self.layout = layout_of(operator.basis)

    def empty_vector(self, dtype=None):
        import torch
        return torch.zeros(self.layout.rows, dtype=dtype or torch.float64, device=self.device)

    def matvec(self, x_local, y_local, mode: Optional[int] = None) -> None:
        """Asynchronous on the library stream (= torch's current stream after ``init_process``)."""
        import torch
        from . import _lib
        assert x_local.numel() == self.layout.rows and y_local.numel() == self.layout.rows
        m = self.mode if mode is None else mode
        if x_local.dtype == torch.complex128:
            status = _lib.lib.ls_b200_dist_matvec_c128(C.byref(self.op._payload), x_local.data_ptr(), y_local.data_ptr(), m)
        else:
            status = _lib.lib.ls_b200_dist_matvec(C.byref(self.op._payload), x_local.data_ptr(), y_local.data_ptr(), m)
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_dist_matvec failed")

    def sync(self) -> None:
        from . import _lib
        _lib.lib.ls_b200_matvec_sync()
        _lib.check_error()

    def dot(self, a_local, b_local):
        """Global <a, b> as a 1-element DEVICE tensor: local dot + one in-place all-reduce on the library's
        communicator and stream -- no host round trip."""
        import torch
        from . import _lib
        s = torch.vdot(a_local, b_local).reshape(1)
        if self.layout.world > 1:
            words = 2 if s.dtype == torch.complex128 else 1
            view = torch.view_as_real(s) if words == 2 else s
            _lib.lib.ls_b200_comm_allreduce_f64(view.data_ptr(), words)
            _lib.check_error()
        return s


# ---- virtual ranks on one GPU -----------------------------------------------------------------------------------
class EmulatedRanks:
    """``world`` virtual ranks of a sharded basis on the current GPU (``ls_b200_emu_build`` /
    ``ls_b200_emu_matvec``): the sharded build, the row balancing, both product forms and the wide index run
    exactly as across processes, with device-to-device copies in place of the NCCL collectives."""

    def __init__(self, make_basis, make_operator, world: int, flags: int = 0, balance: bool = False):
        """``make_basis()`` -> a fresh (unbuilt) Basis, ``make_operator(basis)`` -> an Operator on it; called once per
        virtual rank."""
        from . import _lib
        self.world = world
        self.bases = [make_basis() for _ in range(world)]
        self.ops = [make_operator(b) for b in self.bases]
        bases = (C.c_void_p * world)(*[C.addressof(b._payload) for b in self.bases])
        ops = (C.c_void_p * world)(*[C.addressof(o._payload) for o in self.ops])
        status = _lib.lib.ls_b200_emu_build(bases, ops if balance else None, world, int(flags))
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_emu_build failed")
        self.layouts = [layout_of(b) for b in self.bases]
        self._ops_array = ops

    @classmethod
    def from_model(cls, model, world: int, flags: int = 0, balance: bool = False) -> "EmulatedRanks":
        return cls(model.basis, model.operator, world, flags, balance)

    @property
    def dim(self) -> int:
        return self.layouts[0].dim

    def rebalance(self, costs: np.ndarray) -> bool:
        """``ls_b200_emu_rebalance``: ``costs[r, k]`` = cost of the k-th of 1024 equal pieces of virtual rank r's rows."""
        from . import _lib
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        assert costs.shape == (self.world, 1024)
        bases = (C.c_void_p * self.world)(*[C.addressof(b._payload) for b in self.bases])
        status = _lib.lib.ls_b200_emu_rebalance(bases, self.world, costs.ctypes.data_as(_lib.f64_p))
        _lib.check_error()
        if status < 0:
            raise RuntimeError("ls_b200_emu_rebalance failed")
        for b in self.bases:
            b._keep.pop("views", None)
        self.layouts = [layout_of(b) for b in self.bases]
        return status == 0

    def states(self) -> np.ndarray:
        """The concatenation of the local blocks (== the sorted representatives of the whole basis)."""
        return np.concatenate([np.asarray(b.states) for b in self.bases]) if self.dim else np.zeros(0, np.uint64)

    def matvec(self, x: np.ndarray, mode: int = AUTO) -> np.ndarray:
        """y = H x for a full-length host vector: cut into the ranks' blocks, one product, reassembled."""
        from . import _lib
        cplx = np.iscomplexobj(x)
        dt = np.complex128 if cplx else np.float64
        x = np.ascontiguousarray(x, dtype=dt)
        xs, ys = [], []
        for L in self.layouts:
            xs.append(_lib.DeviceArray.from_numpy(x[L.row_begin:L.row_end]) if L.rows else _lib.DeviceArray(1, dt))
            ys.append(_lib.DeviceArray(max(L.rows, 1), dt))
        xp = (C.c_void_p * self.world)(*[a.ptr for a in xs])
        yp = (C.c_void_p * self.world)(*[a.ptr for a in ys])
        status = _lib.lib.ls_b200_emu_matvec(self._ops_array, self.world, xp, yp, int(mode), 1 if cplx else 0)
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_emu_matvec failed")
        _lib.lib.ls_b200_matvec_sync()
        _lib.check_error()
        out = np.zeros(self.dim, dtype=dt)
        for L, a in zip(self.layouts, ys):
            if L.rows:
                out[L.row_begin:L.row_end] = a.numpy()[:L.rows]
        return out
