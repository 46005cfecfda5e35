"""Multi-GPU sharding of the two hot paths: one process per GPU, ``torch.distributed``
for the plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests).

The reference distributes basis states over Chapel locales by *hash*
(chapel/src/StatesEnumeration.chpl:198-212) and pushes (state, coefficient)
pairs to their owners through GASNet PUTs
(chapel/src/DistributedMatrixVector.chpl:226-339, :545-579).  Here:

* **basis build** -- candidates are addressed by their combinadic index, rank r
  scans the contiguous index range ``shard_bounds(total, P, r)``; the shards are
  already globally ordered, so assembling the basis is one all-gather of the
  counts plus one broadcast of every shard into its slot of the full array.  No
  collective on the data path of the scan itself.
* **matvec** -- rows (= contiguous ranges of sorted representatives) are split
  evenly; the pull-form kernel writes only its own rows, so the single exchange
  step is the all-gather of the x shards (``all_gather_into_tensor`` straight
  into the replicated, padded vector).  y needs no reduction.

The exchange logic below is independent of where the shard was computed: it
takes tensors (CPU or CUDA), so the world_size-2 gloo tests drive it with
shards produced on the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

__all__ = ["shard_bounds", "block_plan", "row_bounds", "balanced_row_bounds", "exchange_shards", "exchange_blocks",
           "ShardedOperator",
           "build_sharded", "init_process"]

ALIGN = 32  # candidate shards start on a multiple of 32 (one bit-sliced word)


def shard_bounds(total: int, world: int, rank: int, align: int = ALIGN) -> Tuple[int, int]:
    """Contiguous candidate-index range of ``rank``: boundaries are multiples of
    ``align``, sizes differ by at most ``align``, the union is [0, total)."""
    words = (total + align - 1) // align
    lo = (words * rank) // world * align
    hi = (words * (rank + 1)) // world * align
    return min(lo, total), min(hi, total)


def block_plan(total: int, world: int, blocks_per_rank: int = 16, min_block: int = 1 << 22,
               align: int = ALIGN) -> List[Tuple[int, int]]:
    """Block-cyclic split of the candidate-index range [0, total): block b = [lo, hi) belongs to rank
    b % world.  Representatives -- and the work of finding them -- are NOT uniform in the candidate
    index: a representative is the smallest member of its orbit, so they crowd into the low indices
    (kagome-36 on two ranks: the lower half holds all 3.15e7 of them and 3/4 of the work).  Dealing
    out ~16 blocks per rank evens that out while every block still yields a sorted range."""
    if total <= 0:
        return []
    size = max(min_block, -(-total // (world * blocks_per_rank)))
    size = -(-size // align) * align
    return [(lo, min(lo + size, total)) for lo in range(0, total, size)]


def row_bounds(dim: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Even row split with a fixed chunk = ceil(dim / world): returns
    (row_begin, row_end, chunk).  Trailing ranks may own fewer (or zero) rows;
    the replicated vector is padded to world * chunk."""
    chunk = (dim + world - 1) // world if dim > 0 else 0
    lo = min(rank * chunk, dim)
    hi = min((rank + 1) * chunk, dim)
    return lo, hi, chunk


def exchange_shards(local, group=None):
    """All ranks contribute a 1-D tensor (their shard, possibly empty); every rank
    receives (full, offsets) where ``full`` is the concatenation in rank order and
    ``offsets[r]`` the start of rank r's shard.  Works on CPU (gloo) and CUDA (NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    count = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, count, group=group)
    counts = counts.cpu().tolist()
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + int(c))
    full = torch.empty(offsets[-1], dtype=local.dtype, device=local.device)
    rank = dist.get_rank(group)
    full[offsets[rank]:offsets[rank + 1]].copy_(local)
    for r in range(world):
        if counts[r] > 0:
            src = dist.get_global_rank(group, r) if group is not None else r
            dist.broadcast(full[offsets[r]:offsets[r + 1]], src=src, group=group)
    return full, offsets


def exchange_blocks(pieces, number_blocks: int, group=None, dtype=None, device=None):
    """Block-cyclic counterpart of ``exchange_shards``: ``pieces`` are this rank's blocks (block
    b = rank + k * world for k = 0, 1, ...) as 1-D tensors, each sorted; every rank receives the
    concatenation of ALL blocks in block order plus the block offsets."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if pieces:
        local = torch.cat(pieces) if len(pieces) > 1 else pieces[0]
    else:
        local = torch.empty(0, dtype=dtype, device=device)
    per_rank = (number_blocks + world - 1) // world
    mine = torch.zeros(max(per_rank, 1), dtype=torch.int64, device=local.device)
    for k, piece in enumerate(pieces):
        mine[k] = piece.numel()
    counts = torch.zeros(world * max(per_rank, 1), dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.cpu().view(world, max(per_rank, 1))
    full, rank_offsets = exchange_shards(local, group)
    out = torch.empty_like(full)
    block_offsets = [0]
    within = [0] * world
    for b in range(number_blocks):
        r, k = b % world, b // world
        n = int(counts[r, k])
        src = rank_offsets[r] + within[r]
        out[block_offsets[-1]:block_offsets[-1] + n].copy_(full[src:src + n])
        within[r] += n
        block_offsets.append(block_offsets[-1] + n)
    return out, block_offsets


# ---- CUDA side ---------------------------------------------------------------------------
class _RawDevice:
    """Zero-copy view of a raw device pointer for ``torch.as_tensor``."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}


def tensor_from_pointer(ptr: int, count: int, dtype: str):
    """dtype: 'u8' (uint64 viewed as int64, NCCL has no uint64 arithmetic but moves bytes) or 'f8'."""
    import torch
    if count == 0:
        return torch.empty(0, dtype=torch.int64 if dtype == "u8" else torch.float64, device="cuda")
    typestr = "<i8" if dtype == "u8" else "<f8"
    return torch.as_tensor(_RawDevice(ptr, count, typestr), device="cuda")


_stream = None


def init_process(local_rank: Optional[int] = None):
    """Select this rank's GPU and order all library work on a dedicated torch
    stream, so that NCCL collectives and our kernels interleave without host syncs."""
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


def build_sharded(basis, group=None) -> List[int]:
    """Build ``basis`` across the ranks of ``group``: every rank scans its blocks of the candidate
    range (``block_plan``: block-cyclic, because the work is concentrated at low indices) on its own
    GPU, the pieces are exchanged over NCCL and the full sorted representative list (+ norms) is
    installed on every rank.  Returns the block offsets into the representative list."""
    import torch
    import torch.distributed as dist
    from . import _lib
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    total = basis.number_candidates
    plan = block_plan(total, world)
    reps, norms = [], []
    with_norms = basis.has_permutation_symmetries
    mine = plan[rank::world]
    d_reps = d_norms = 0
    if mine:
        size = plan[0][1] - plan[0][0]
        d_reps, d_norms, counts = basis.build_blocks(mine[0][0], size, size * world, len(mine))
        all_reps = tensor_from_pointer(d_reps, sum(counts), "u8")
        all_norms = tensor_from_pointer(d_norms, sum(counts), "f8") if with_norms else None
        start = 0
        for c in counts:
            reps.append(all_reps[start:start + c])
            if with_norms:
                norms.append(all_norms[start:start + c])
            start += c
    raw = [(d_reps, d_norms)]
    full_reps, offsets = exchange_blocks(reps, len(plan), group, dtype=torch.int64, device="cuda")
    dim = offsets[-1]
    # the library takes ownership of buffers it allocated itself
    own_reps = _lib.lib.ls_b200_device_malloc(max(8 * dim, 8))
    tensor_from_pointer(own_reps, dim, "u8").copy_(full_reps)
    own_norms = None
    if with_norms:
        full_norms, _ = exchange_blocks(norms, len(plan), group, dtype=torch.float64, device="cuda")
        own_norms = _lib.lib.ls_b200_device_malloc(max(8 * dim, 8))
        tensor_from_pointer(own_norms, dim, "f8").copy_(full_norms)
    torch.cuda.current_stream().synchronize()
    for d_reps, d_norms in raw:
        if d_reps:
            _lib.lib.ls_b200_device_free(d_reps)
        if d_norms:
            _lib.lib.ls_b200_device_free(d_norms)
    basis.set_representatives_device(own_reps, own_norms, dim)
    return offsets


@dataclass
class _Layout:
    dim: int
    world: int
    rank: int
    row_begin: int
    row_end: int
    chunk: int                      # rows of the largest shard: the all-gather slot size
    bounds: Optional[list] = None   # (row_begin, row_end) of every rank; None: the even split of row_bounds


def balanced_row_bounds(block_costs, block_rows: int, dim: int, world: int):
    """Contiguous row ranges of (nearly) equal cost from per-block costs (block b = rows
    [b block_rows, (b + 1) block_rows)): rank r ends at the first block boundary where the running
    cost reaches (r + 1) / world of the total.  Rows of the sorted basis do not cost the same -- on
    kagome-36 the first eighth of the rows holds 0.87x, the last 1.12x the mean number of matrix
    elements -- and a step is as slow as its slowest rank."""
    total = float(sum(block_costs))
    bounds, lo, acc, b = [], 0, 0.0, 0
    for r in range(world):
        target = total * (r + 1) / world
        while b < len(block_costs) and (acc + block_costs[b] <= target or r == world - 1):
            acc += block_costs[b]
            b += 1
        # take the block that straddles the target if that lands closer to it
        if r < world - 1 and b < len(block_costs) and target - acc > acc + block_costs[b] - target:
            acc += block_costs[b]
            b += 1
        hi = dim if r == world - 1 else min(dim, b * block_rows)
        bounds.append((lo, max(lo, hi)))
        lo = max(lo, hi)
    return bounds


class ShardedOperator:
    """y = H x with rows sharded over the ranks and x replicated by all-gather.

    ``x_full`` is a padded replicated vector of length world * chunk; ``matvec``
    computes this rank's rows into its slot of ``y_full`` and all-gathers in
    place, so the output can be fed straight back in (Lanczos)."""

    device = "cuda"

    def __init__(self, operator, group=None, dim: Optional[int] = None, bounds=None, balance: bool = True):
        """``bounds``: explicit (row_begin, row_end) per rank; otherwise, with an operator and ``balance``, the
        rows are split by matrix-element count (every rank computes the same split from the replicated basis),
        else evenly."""
        import torch.distributed as dist
        self.op = operator
        self.group = group
        if dim is None:
            dim = operator.basis.number_states
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if bounds is None and balance and operator is not None and world > 1 and dim >= 4096 * world:
            blocks = 64 * world
            block_rows = -(-dim // blocks)
            costs = []
            for b in range(blocks):
                lo_b, hi_b = min(b * block_rows, dim), min((b + 1) * block_rows, dim)
                # 4 row-proportional units (alpha, norm, x, y traffic and the term scan) per row on top of its elements
                costs.append(operator.count_matrix_elements(lo_b, hi_b) + 4 * (hi_b - lo_b) if hi_b > lo_b else 0)
            bounds = balanced_row_bounds(costs, block_rows, dim, world)
        if bounds is not None:
            bounds = [(int(a), int(b)) for a, b in bounds]
            assert len(bounds) == world and bounds[0][0] == 0 and bounds[-1][1] == dim
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            chunk = max(hi - lo for lo, hi in bounds)
            even = [row_bounds(dim, world, r)[:2] for r in range(world)]
            if [tuple(b) for b in bounds] == [tuple(e) for e in even]:
                bounds = None
        if bounds is None:
            lo, hi, chunk = row_bounds(dim, world, rank)
        else:
            lo, hi = bounds[rank]
        self.layout = _Layout(dim, world, rank, lo, hi, chunk, bounds)
        self._gather_buffers = {}

    def empty_vector(self, dtype=None):
        import torch
        L = self.layout
        return torch.zeros(L.world * L.chunk, dtype=dtype or torch.float64, device=self.device)

    def local_rows(self, v):
        L = self.layout
        return v[L.row_begin:L.row_end]

    def matvec(self, x_full, y_full, gather: bool = True) -> None:
        import torch
        import torch.distributed as dist
        L = self.layout
        cplx = x_full.dtype == torch.complex128
        if L.row_end > L.row_begin:
            self._local_rows(x_full, y_full, L.row_begin, L.row_end, cplx)
        if gather and L.world > 1:
            # While NCCL replicates y: canonicalise the matrix elements for the NEXT product (they do not depend
            # on the vector).  Every product still runs both phases exactly once.
            self.gather_rows(y_full, overlap=(lambda: self._prepare_rows(L.row_begin, L.row_end, cplx))
                             if L.row_end > L.row_begin else None)

    def gather_rows(self, v_full, overlap=None) -> None:
        """Replicate every rank's rows of ``v_full`` on all ranks (one all-gather); ``overlap`` is called while
        the collective is in flight."""
        import torch
        import torch.distributed as dist
        L = self.layout
        if L.world == 1:
            return
        real = (lambda t: torch.view_as_real(t)) if v_full.dtype == torch.complex128 else (lambda t: t)
        if L.bounds is None:
            # even split: rank r's rows already sit in slot r of the padded vector -- gather in place
            mine = v_full[L.rank * L.chunk:(L.rank + 1) * L.chunk]
            work = dist.all_gather_into_tensor(real(v_full), real(mine), group=self.group, async_op=True)
            if overlap is not None:
                overlap()
            work.wait()
            return
        # balanced (uneven) split: gather fixed-size slots into a side buffer, then copy the other ranks' rows home
        key = (v_full.dtype, v_full.device)
        buf = self._gather_buffers.get(key)
        if buf is None:
            buf = torch.zeros(L.world * L.chunk, dtype=v_full.dtype, device=v_full.device)
            self._gather_buffers[key] = buf
        mine = buf[L.rank * L.chunk:(L.rank + 1) * L.chunk]
        mine[:L.row_end - L.row_begin].copy_(v_full[L.row_begin:L.row_end])
        work = dist.all_gather_into_tensor(real(buf), real(mine), group=self.group, async_op=True)
        if overlap is not None:
            overlap()
        work.wait()
        for r, (lo, hi) in enumerate(L.bounds):
            if r != L.rank and hi > lo:
                v_full[lo:hi].copy_(buf[r * L.chunk:r * L.chunk + (hi - lo)])

    def _local_rows(self, x_full, y_full, row_begin: int, row_end: int, cplx: bool) -> None:
        """y_full[row_begin:row_end] = (H x)[row_begin:row_end] on this rank's GPU."""
        y_ptr = y_full.data_ptr() + row_begin * y_full.element_size()
        if self.layout.world > 1:
            # phase 2 (phase 1 ran during the previous all-gather, or runs now on the first call)
            self.op.matvec_device_phase(2, x_full.data_ptr(), y_ptr, row_begin, row_end, complex_vectors=cplx)
        else:
            self.op.matvec_device(x_full.data_ptr(), y_ptr, row_begin, row_end, complex_vectors=cplx)

    def _prepare_rows(self, row_begin: int, row_end: int, cplx: bool) -> None:
        """Work of the next product that does not depend on the vector (overridable; no-op without an operator)."""
        if self.op is not None:
            self.op.matvec_device_phase(1, 0, 0, row_begin, row_end, complex_vectors=cplx)

    def dot(self, a_full, b_full):
        """Global <a, b> from the local rows (one all-reduce of a scalar)."""
        import torch
        import torch.distributed as dist
        s = torch.vdot(self.local_rows(a_full), self.local_rows(b_full)).reshape(1)
        if self.layout.world > 1:
            if s.dtype == torch.complex128:
                dist.all_reduce(torch.view_as_real(s), group=self.group)
            else:
                dist.all_reduce(s, group=self.group)
        return s
