"""CPU tests of the multi-GPU host logic -- the planning functions the library's sharded build runs
(csrc/dist.cu, exported through the C ABI as ls_b200_plan_*), single-process simulations of whole exchanges, and
world_size-2/3 gloo runs in which ``torch.distributed.all_to_all_single`` stands in for the grouped NCCL
send/recv and the CPU oracle for the per-rank scan: block-cyclic pieces must end up as the contiguous sorted row
ranges the plan promises, with norms, and the owner routing of the push form must deliver every record to the
rank that holds its row."""
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

from lattice_symmetries_b200.distributed import (  # noqa: E402
    balanced_bounds, cyclic_share, even_bounds, plan_blocks, plan_redistribution, rebalance_bounds, row_bounds,
    shard_bounds)


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


@pytest.mark.parametrize("total", [0, 1, 31, 4096, 2704156, 4537567650, 269128937220])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_block_plan_partition(total, world):
    """Blocks tile [0, total), are equal (a power of two >= 2^20 candidates; the tail may be short) and numerous
    enough to deal every density regime out evenly, yet at most ~8192 per rank."""
    plan = plan_blocks(total, world)
    assert sum(hi - lo for lo, hi in plan) == total
    prev = 0
    for lo, hi in plan:
        assert lo == prev and lo % 32 == 0 and hi > lo
        prev = hi
    assert prev == total
    sizes = [hi - lo for lo, hi in plan]
    if sizes:
        block = sizes[0] if len(sizes) > 1 else None
        if block is not None:
            assert block >= 1 << 20 and block & (block - 1) == 0
            assert all(s == block for s in sizes[:-1]) and sizes[-1] <= block
        assert len(plan) <= 8192 * world + 1
    # fine enough to deal out: a long range gives every rank many blocks
    if total > (1 << 20) * 64 * world:
        assert len(plan) >= 32 * world


@pytest.mark.parametrize("dim", [0, 1, 5, 13, 28968])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_row_bounds_partition(dim, world):
    prev = 0
    for r in range(world):
        lo, hi, chunk = row_bounds(dim, world, r)
        assert lo == prev and hi - lo <= chunk
        prev = hi
    assert prev == dim
    b = even_bounds(dim, world)
    assert b[0] == 0 and b[-1] == dim and all(x <= y for x, y in zip(b, b[1:]))
    assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_balanced_bounds(world):
    rng = np.random.default_rng(3)
    blocks = 64 * world
    edges = np.concatenate([[0], np.cumsum(rng.integers(1, 200, size=blocks))])
    costs = rng.uniform(0.5, 1.5, size=blocks) * np.linspace(0.8, 1.2, blocks)
    bounds = balanced_bounds(edges, costs, world)
    assert bounds[0] == 0 and bounds[-1] == edges[-1] and len(bounds) == world + 1
    assert all(a <= b for a, b in zip(bounds, bounds[1:]))
    assert set(bounds) <= set(int(e) for e in edges)
    # every rank within one block's cost of the mean
    cum = np.concatenate([[0], np.cumsum(costs)])
    per = [cum[list(edges).index(b)] for b in bounds]
    shares = np.diff(per)
    assert np.all(np.abs(shares - costs.sum() / world) <= costs.max() + 1e-9)


@pytest.mark.parametrize("total", [1, 31, 32, 4096, 2704156, 4537567650, 269128937220])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_cyclic_shares_partition_the_candidates(total, world):
    """The block-cyclic scan of the sharded build (what every rank walks as one virtual range): the ranks' shares tile
    the candidate range exactly -- also when the last block is partial or there are fewer blocks than ranks."""
    plan = plan_blocks(total, world)
    shares = [cyclic_share(total, world, r) for r in range(world)]
    shift, blocks_total = shares[0][0], shares[0][1]
    block = 32 << shift
    assert all(s[0] == shift and s[1] == blocks_total for s in shares)
    assert blocks_total == len(plan) == -(-total // block)
    assert sum(s[2] for s in shares) == blocks_total
    assert sum(s[3] for s in shares) == total
    for r, (_, _, mine, candidates) in enumerate(shares):
        want_blocks = plan[r::world]
        assert mine == len(want_blocks)
        assert candidates == sum(hi - lo for lo, hi in want_blocks)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rebalance_bounds_equal_cost_under_a_row_cap(world):
    """Measured cost per row rising along the sorted list (the kagome-42 profile: 2.6x from the first to the last rank):
    without a cap every rank gets the same cost; with the cap the first ranks stop at the cap and the REST is balanced."""
    rng = np.random.default_rng(5)
    dim = 1_000_000
    blocks = 64 * world
    edges = np.concatenate([[0], np.sort(rng.choice(np.arange(1, dim), size=blocks - 1, replace=False)), [dim]])
    density = lambda g: 1.0 + 1.6 * g / dim
    mid = 0.5 * (edges[:-1] + edges[1:])
    costs = density(mid) * np.diff(edges)
    cum = lambda g: g + 0.8 * g * g / dim
    free = rebalance_bounds(edges, costs, world, dim)
    assert free[0] == 0 and free[-1] == dim and all(a <= b for a, b in zip(free, free[1:]))
    shares = np.diff([cum(b) for b in free])
    assert np.all(np.abs(shares - cum(dim) / world) <= 0.01 * cum(dim) / world)
    cap = int(dim / world * 1.1)
    capped = rebalance_bounds(edges, costs, world, cap)
    rows = np.diff(capped)
    assert capped[0] == 0 and capped[-1] == dim and rows.max() <= cap and rows[0] == cap
    hit = int(np.sum(rows == cap))
    assert 1 <= hit < world
    rest = np.diff([cum(b) for b in capped[hit:]])          # the ranks below the cap share the remaining cost evenly
    assert np.all(np.abs(rest - rest.mean()) <= 0.02 * rest.mean())
    # a cap that cannot hold the basis at all: every rank takes the cap, the last one the remainder
    tight = rebalance_bounds(edges, costs, world, dim // world - 1)
    assert tight[-1] == dim and np.all(np.diff(tight)[:-1] <= dim // world - 1)


def test_bounds_properties():
    """Any blocks, any costs (empty blocks, zero costs, all of the cost in one block, costs FALLING along the list): the
    bounds are monotone, cover [0, dim], and a feasible row cap -- it stands for what a rank's memory holds -- is honoured
    by every rank, the last ones included."""
    from hypothesis import given, settings, strategies as st

    @st.composite
    def blocks(draw):
        nb = draw(st.integers(1, 40))
        lens = draw(st.lists(st.integers(0, 1000), min_size=nb, max_size=nb))
        costs = draw(st.lists(st.floats(0, 1e6, allow_nan=False), min_size=nb, max_size=nb))
        return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), np.array(costs), draw(st.integers(1, 9))

    @settings(max_examples=400, deadline=None, derandomize=True)
    @given(blocks(), st.integers(1, 5000))
    def check(b, cap):
        edges, costs, world = b
        dim = int(edges[-1])
        for out in (balanced_bounds(edges, costs, world), rebalance_bounds(edges, costs, world, cap)):
            assert len(out) == world + 1 and out[0] == 0 and out[-1] == dim
            assert all(a <= c for a, c in zip(out, out[1:]))
        if cap * world >= dim > 0:
            assert np.diff(rebalance_bounds(edges, costs, world, cap)).max() <= cap

    check()
    # the case the forward pass alone missed: cost falling along the list leaves the LAST rank over the cap
    edges = np.arange(0, 101, 10)
    costs = np.linspace(10.0, 1.0, 10)
    out = rebalance_bounds(edges, costs, 2, 60)
    assert out == [0, 40, 100]
    assert rebalance_bounds(edges, costs, 2, 100)[1] < 40     # (uncapped, rank 0 takes fewer, dearer rows)
    assert rebalance_bounds(np.array([0, 0, 3]), np.array([1.0, 0.0]), 2, 2) == [0, 1, 3]


def _simulate(world: int, lengths, owners, bounds, data):
    """All ranks in one process: what every rank ends up with after the planned all-to-all-v."""
    plans = [plan_redistribution(world, r, lengths, owners, bounds) for r in range(world)]
    # what each rank holds now: its pieces, concatenated in piece order
    starts = np.concatenate([[0], np.cumsum(lengths)])
    held = [np.concatenate([data[starts[i]:starts[i + 1]] for i in range(len(lengths)) if owners[i] == r] +
                           [np.zeros(0, dtype=data.dtype)]) for r in range(world)]
    out = []
    for r in range(world):
        p = plans[r]
        assert int(p.scount.sum()) == held[r].shape[0]
        recv = np.zeros(int(p.rcount.sum()), dtype=data.dtype)
        for s in range(world):
            q = plans[s]
            assert q.scount[r] == p.rcount[s]
            recv[p.rdispl[s]:p.rdispl[s] + p.rcount[s]] = held[s][q.sdispl[r]:q.sdispl[r] + q.scount[r]]
        mine = np.zeros(bounds[r + 1] - bounds[r], dtype=data.dtype)
        covered = 0
        for src, dst, n in p.places:
            mine[dst:dst + n] = recv[src:src + n]
            covered += n
        assert covered == mine.shape[0]
        out.append(mine)
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plan_redistribution_simulated(world, seed):
    """Block-cyclic pieces of random (also empty) lengths -> contiguous ranges, even and uneven targets."""
    rng = np.random.default_rng(seed)
    nb = int(rng.integers(1, 40))
    lengths = rng.integers(0, 50, size=nb)
    lengths[rng.integers(0, nb)] = 0
    owners = np.arange(nb) % world
    dim = int(lengths.sum())
    data = np.arange(dim, dtype=np.int64) * 3 + 7
    for bounds in (even_bounds(dim, world),
                   sorted([0, dim] + [int(v) for v in rng.integers(0, dim + 1, size=world - 1)])):
        parts = _simulate(world, lengths, owners, bounds, data)
        assert np.array_equal(np.concatenate(parts), data)
        for r in range(world):
            assert parts[r].shape[0] == bounds[r + 1] - bounds[r]
    # second use: contiguous ranges -> other contiguous ranges (the balancing step)
    old = even_bounds(dim, world)
    new = sorted([0, dim] + [int(v) for v in rng.integers(0, dim + 1, size=world - 1)])
    parts = _simulate(world, np.diff(old), np.arange(world), new, data)
    assert np.array_equal(np.concatenate(parts), data)


def _worker_build(rank, world, port, out):
    """Sharded build under gloo: the oracle scans this rank's blocks of plan_blocks, one all_to_all_single per
    array moves every piece to the rank that owns its rows (the plan of csrc/dist.cu)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ls_oracle as oracle
        from lattice_symmetries_b200 import lattices
        import helpers as H
        model = lattices.heisenberg_chain(16)
        p = H.Problem(model.name, model.number_sites, model.expression, hamming_weight=model.hamming_weight,
                      spin_inversion=model.spin_inversion, symmetries=model.symmetries)
        ob = p.oracle_basis(oracle)
        lib = oracle.lib()
        lo, hi = ob.min_state(), ob.max_state()
        r_lo = int(lib.oracle_fixed_hamming_state_to_index(lo))
        total = int(lib.oracle_fixed_hamming_state_to_index(hi)) - r_lo + 1
        # (the library's plan starts at 2^20 candidates per block; cut finer here so that a 16-site chain has pieces)
        step = 32 * 7
        plan = [(a, min(total, a + step)) for a in range(0, total, step)]
        hw = model.hamming_weight
        mine_reps, mine_norms, counts = [], [], []
        for b in range(rank, len(plan), world):
            a, e = plan[b]
            first = int(lib.oracle_fixed_hamming_index_to_state(r_lo + a, hw))
            last = int(lib.oracle_fixed_hamming_index_to_state(r_lo + e - 1, hw))
            reps = ob.enumerate_range(first, last)
            mine_reps.append(reps.astype(np.int64))
            mine_norms.append(ob.group.state_info(reps)[2])
            counts.append(reps.shape[0])
        per_rank = (len(plan) + world - 1) // world
        mine = torch.zeros(per_rank, dtype=torch.int64)
        mine[:len(counts)] = torch.tensor(counts, dtype=torch.int64)
        table = torch.zeros(world * per_rank, dtype=torch.int64)
        dist.all_gather_into_tensor(table, mine)
        table = table.view(world, per_rank).numpy()
        lengths = np.array([table[b % world, b // world] for b in range(len(plan))])
        owners = np.arange(len(plan)) % world
        dim = int(lengths.sum())
        bounds = even_bounds(dim, world)
        pl = plan_redistribution(world, rank, lengths, owners, bounds)

        def exchange(pieces, dtype):
            send = torch.from_numpy(np.concatenate(pieces + [np.zeros(0, dtype=dtype)]))
            recv = torch.empty(int(pl.rcount.sum()), dtype=send.dtype)
            dist.all_to_all_single(recv, send, [int(c) for c in pl.rcount], [int(c) for c in pl.scount])
            local = np.zeros(bounds[rank + 1] - bounds[rank], dtype=dtype)
            for src, dst, n in pl.places:
                local[dst:dst + n] = recv.numpy()[src:src + n]
            return local

        reps_local = exchange(mine_reps, np.int64).astype(np.uint64)
        norms_local = exchange(mine_norms, np.float64)
        want = ob.enumerate()
        want_norms = ob.group.state_info(want)[2]
        ok = np.array_equal(reps_local, want[bounds[rank]:bounds[rank + 1]]) and np.array_equal(
            norms_local.view(np.uint64), want_norms[bounds[rank]:bounds[rank + 1]].view(np.uint64))
        # owner routing of the push form: every representative belongs to the rank whose range holds it
        firsts = torch.zeros(world, dtype=torch.int64)
        mine_first = torch.tensor([int(reps_local[0]) if reps_local.size else np.iinfo(np.int64).max], dtype=torch.int64)
        dist.all_gather_into_tensor(firsts, mine_first)
        splitters = firsts.numpy().astype(np.uint64)
        owner = np.zeros(want.shape[0], dtype=np.int64)
        for k in range(1, world):
            owner += (splitters[k] <= want).astype(np.int64)
        ok = ok and np.array_equal(owner, np.searchsorted(np.array(bounds[1:]), np.arange(want.shape[0]), side="right"))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _worker_push(rank, world, port, out):
    """The all-to-all product under gloo, the oracle standing in for the kernels: every rank applies the operator to
    ITS columns, routes (representative, coefficient) records to the owners by splitter search, the owner ranks
    them locally and accumulates -- the result must be the rows of the oracle's full product."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ls_oracle as oracle
        from lattice_symmetries_b200 import lattices
        import helpers as H
        model = lattices.kagome_heisenberg(12)
        p = H.Problem(model.name, model.number_sites, model.expression, hamming_weight=model.hamming_weight,
                      spin_inversion=model.spin_inversion, symmetries=model.symmetries)
        ob, reps, index, off, diag = p.oracle_setup(oracle)
        dim = reps.shape[0]
        bounds = even_bounds(dim, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        rng = np.random.default_rng(11)
        x = rng.standard_normal(dim)
        norms = ob.group.state_info(reps)[2]
        # columns of this rank: H|alpha_i> = sum c |beta>, canonicalise, coefficient chi c n_beta / n_alpha x_i
        betas, coeffs, offsets = oracle.apply_off_diag(off, reps[lo:hi])
        rep_b, chi, n_b = ob.group.state_info(betas)
        col = np.repeat(np.arange(lo, hi), np.diff(offsets))
        c = (chi * coeffs).real * n_b / norms[col] * x[col]
        keep = c != 0
        rep_b, c = rep_b[keep], c[keep]
        splitters = reps[np.minimum(np.array(bounds[:-1]), dim - 1)]
        owner = np.zeros(rep_b.shape[0], dtype=np.int64)
        for k in range(1, world):
            owner += (splitters[k] <= rep_b).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        scount = np.bincount(owner, minlength=world)
        cnt = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(cnt, torch.from_numpy(scount.astype(np.int64)))
        rcount = [int(v) for v in cnt]

        def a2a(values):
            send = torch.from_numpy(np.ascontiguousarray(values[order]))
            recv = torch.empty(sum(rcount), dtype=send.dtype)
            dist.all_to_all_single(recv, send, rcount, [int(v) for v in scount])
            return recv.numpy()

        got_rep = a2a(rep_b.astype(np.int64)).astype(np.uint64)
        got_c = a2a(c)
        local_index = oracle.Index(reps[lo:hi], ob.number_bits, 22)
        j = local_index(got_rep)
        y = oracle.apply_diag(diag, reps[lo:hi], x[lo:hi]) if diag.n else np.zeros(hi - lo)
        assert np.all(j >= 0)
        np.add.at(y, j, got_c)   # (n_beta is already in c: the oracle's state_info returned it per beta)
        want, _ = oracle.matvec(ob, off, diag, index, x)
        err = np.linalg.norm(y - want[lo:hi]) / max(np.linalg.norm(want), 1e-300)
        out[rank] = bool(err < 1e-12)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker_build, _worker_push])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo(worker, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        out = manager.dict()
        port = _free_port()
        procs = [ctx.Process(target=worker, args=(r, world, port, out)) for r in range(world)]
        for pr in procs:
            pr.start()
        for pr in procs:
            pr.join(timeout=300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        assert dict(out) == {r: True for r in range(world)}
