"""world_size-2/3 gloo tests of the multi-GPU host logic: candidate
shards concatenate to the sorted basis, the padded row layout all-gathers
correctly, global dots reduce.  The per-rank compute is stood in for by the CPU
oracle / a dense matrix; the exchange code is the one the NCCL path runs."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from lattice_symmetries_b200.distributed import balanced_row_bounds, block_plan, row_bounds, shard_bounds  # noqa: E402


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [0, 1, 31, 32, 33, 1000, 2704156, 9075135300])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_bounds_partition(total, world):
    prev = 0
    for r in range(world):
        lo, hi = shard_bounds(total, world, r)
        assert lo == prev and lo <= hi <= total
        assert lo % 32 == 0 or lo == total
        prev = hi
    assert prev == total
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 64


@pytest.mark.parametrize("total", [0, 1, 31, 4096, 2704156, 4537567650])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_block_plan_partition(total, world):
    plan = block_plan(total, world)
    assert sum(hi - lo for lo, hi in plan) == total
    prev = 0
    for lo, hi in plan:
        assert lo == prev and lo % 32 == 0 and hi > lo
        prev = hi
    assert prev == total
    if total > (1 << 22) * world * 16:
        assert world * 16 <= len(plan) <= world * 16 + 1


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_balanced_row_bounds(world):
    rng = np.random.default_rng(3)
    blocks, block_rows = 64 * world, 100
    dim = blocks * block_rows - 37
    costs = list(np.linspace(0.8, 1.2, blocks) * 1000 + rng.integers(0, 50, blocks))
    bounds = balanced_row_bounds(costs, block_rows, dim, world)
    assert len(bounds) == world and bounds[0][0] == 0 and bounds[-1][1] == dim
    assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    per_rank = [sum(costs[lo // block_rows:-(-hi // block_rows)]) for lo, hi in bounds]
    assert max(per_rank) <= 1.03 * sum(costs) / world + max(costs)
    # rows no longer split evenly: later (costlier) ranks own fewer rows
    if world > 1:
        assert bounds[0][1] - bounds[0][0] > bounds[-1][1] - bounds[-1][0]


@pytest.mark.parametrize("dim", [0, 1, 5, 13, 28968])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_row_bounds_partition(dim, world):
    prev = 0
    for r in range(world):
        lo, hi, chunk = row_bounds(dim, world, r)
        assert lo == min(r * chunk, dim) and lo <= hi <= dim and hi - lo <= chunk
        prev = max(prev, hi)
    assert prev == dim
    assert world * row_bounds(dim, world, 0)[2] >= dim


def _worker_build(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import helpers as H
        from lattice_symmetries_b200 import lattices as L
        from lattice_symmetries_b200.distributed import exchange_shards
        from oracle import ls_oracle as oracle
        m = L.heisenberg_chain(16)
        p = H.Problem(m.name, m.number_sites, m.expression, hamming_weight=m.hamming_weight,
                      spin_inversion=m.spin_inversion, symmetries=m.symmetries)
        ob = p.oracle_basis(oracle)
        lib = oracle.lib()
        r_lo = int(lib.oracle_fixed_hamming_state_to_index(ob.min_state()))
        r_hi = int(lib.oracle_fixed_hamming_state_to_index(ob.max_state()))
        total = r_hi - r_lo + 1
        lo, hi = shard_bounds(total, world, rank)
        if hi > lo:
            first = int(lib.oracle_fixed_hamming_index_to_state(r_lo + lo, m.hamming_weight))
            last = int(lib.oracle_fixed_hamming_index_to_state(r_lo + hi - 1, m.hamming_weight))
            local = ob.enumerate_range(first, last)
        else:
            local = np.zeros(0, dtype=np.uint64)
        full, offsets = exchange_shards(torch.from_numpy(local.view(np.int64)))
        want = ob.enumerate()
        ok = np.array_equal(full.numpy().view(np.uint64), want) and offsets[-1] == want.shape[0]
        # norms ride the same exchange
        norms = ob.group.state_info(local)[2]
        full_norms, _ = exchange_shards(torch.from_numpy(norms))
        ok = ok and np.array_equal(full_norms.numpy(), ob.group.state_info(want)[2])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _worker_build_blocks(rank, world, port, out):
    """Block-cyclic build (distributed.build_sharded's plan) with the oracle standing in for the GPU."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import helpers as H
        from lattice_symmetries_b200 import lattices as L
        from lattice_symmetries_b200.distributed import exchange_blocks
        from oracle import ls_oracle as oracle
        m = L.heisenberg_chain(16)
        p = H.Problem(m.name, m.number_sites, m.expression, hamming_weight=m.hamming_weight,
                      spin_inversion=m.spin_inversion, symmetries=m.symmetries)
        ob = p.oracle_basis(oracle)
        lib = oracle.lib()
        r_lo = int(lib.oracle_fixed_hamming_state_to_index(ob.min_state()))
        r_hi = int(lib.oracle_fixed_hamming_state_to_index(ob.max_state()))
        total = r_hi - r_lo + 1
        plan = block_plan(total, world, blocks_per_rank=3, min_block=64)
        assert len(plan) >= 2 * world
        pieces, norms = [], []
        for lo, hi in plan[rank::world]:
            first = int(lib.oracle_fixed_hamming_index_to_state(r_lo + lo, m.hamming_weight))
            last = int(lib.oracle_fixed_hamming_index_to_state(r_lo + hi - 1, m.hamming_weight))
            local = ob.enumerate_range(first, last)
            pieces.append(torch.from_numpy(local.view(np.int64)))
            norms.append(torch.from_numpy(ob.group.state_info(local)[2]))
        full, offsets = exchange_blocks(pieces, len(plan), dtype=torch.int64, device="cpu")
        want = ob.enumerate()
        ok = np.array_equal(full.numpy().view(np.uint64), want) and offsets[-1] == want.shape[0]
        ok = ok and len(offsets) == len(plan) + 1
        full_norms, _ = exchange_blocks(norms, len(plan), dtype=torch.float64, device="cpu")
        ok = ok and np.array_equal(full_norms.numpy(), ob.group.state_info(want)[2])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _worker_matvec(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lattice_symmetries_b200.distributed import ShardedOperator

        dim = 37
        rng = np.random.default_rng(0)
        A = rng.standard_normal((dim, dim))
        A = A + A.T

        class Dense(ShardedOperator):
            device = "cpu"

            def _local_rows(self, x_full, y_full, row_begin, row_end, cplx):
                y_full[row_begin:row_end] = torch.from_numpy(A[row_begin:row_end] @ x_full[:dim].numpy())

        # uneven (balanced-style) shards: same answers through the side-buffer gather
        cuts = [0] + sorted({(dim * (2 * r + 1)) // (2 * world + 1) for r in range(1, world)}) + [dim]
        while len(cuts) < world + 1:
            cuts.insert(-1, cuts[-2])
        uneven = Dense(None, dim=dim, bounds=list(zip(cuts[:-1], cuts[1:])))
        xu = uneven.empty_vector()
        xu[:dim] = torch.from_numpy(np.arange(dim, dtype=np.float64))
        yu = uneven.empty_vector()
        uneven.matvec(xu, yu)
        assert np.allclose(yu[:dim].numpy(), A @ np.arange(dim, dtype=np.float64))
        sh = Dense(None, dim=dim)
        x = sh.empty_vector()
        x[:dim] = torch.from_numpy(rng.standard_normal(dim))
        y = sh.empty_vector()
        sh.matvec(x, y)
        ok = np.allclose(y[:dim].numpy(), A @ x[:dim].numpy()) and float(y[dim:].abs().sum()) == 0.0
        # feed the gathered output straight back in (Lanczos-style) and check the global dot
        z = sh.empty_vector()
        sh.matvec(y, z)
        ok = ok and np.allclose(z[:dim].numpy(), A @ (A @ x[:dim].numpy()))
        d = float(sh.dot(y, z).item())
        ok = ok and np.isclose(d, float(y[:dim].numpy() @ z[:dim].numpy()))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker_build, _worker_build_blocks, _worker_matvec])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo(worker, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
            assert p.exitcode == 0
        assert dict(out) == {r: True for r in range(world)}
