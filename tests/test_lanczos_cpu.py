"""CPU tests of the Lanczos drivers on a stand-in operator: the recurrence, the stopping rules and the
checkpoint / resume logic are host code and do not care where the vectors live.  (On the GPU the same drivers run on
the library's products: tests/test_gpu_parity.py, tests/nccl_worker.py.)"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from lattice_symmetries_b200.distributed import Layout
from lattice_symmetries_b200.lanczos import lanczos_block_thick_restart, lanczos_ground_state, lanczos_thick_restart


class DenseOperator:
    """The interface the solvers use (lanczos._wrap), backed by a dense symmetric / Hermitian matrix on the CPU."""
    device = "cpu"

    def __init__(self, matrix: np.ndarray):
        self.H = torch.from_numpy(matrix)
        dim = matrix.shape[0]
        self.layout = Layout(1, 0, dim, 0, dim, 0, 0, 0, [0, dim])
        self.products = 0

    def empty_vector(self, dtype=None):
        return torch.zeros(self.layout.dim, dtype=dtype or torch.float64)

    def matvec(self, x, y, mode=None):
        self.products += 1
        y.copy_(self.H.to(x.dtype) @ x)

    def dot(self, a, b):
        return torch.vdot(a, b).reshape(1)

    def sync(self):
        pass


class BlockDenseOperator(DenseOperator):
    """... with the block product of the library (one pass over the matrix for all right-hand sides)."""

    def __init__(self, matrix):
        super().__init__(matrix)
        self.block_products = 0

    def matvec_block(self, X, Y):
        assert X.dtype == torch.float64 and X.is_contiguous() and Y.is_contiguous() and X.shape == Y.shape
        self.block_products += 1
        Y.copy_(X @ self.H.t())


def _matrix(dim, seed, cplx=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((dim, dim)) + (1j * rng.standard_normal((dim, dim)) if cplx else 0)
    h = (a + a.conj().T) / 2
    return h + np.diag(np.linspace(-dim / 4, dim / 4, dim))   # spread the spectrum: Lanczos converges in ~100 steps


def test_ground_state_and_eigenvector():
    H = _matrix(400, 1)
    w, v = np.linalg.eigh(H)
    op = DenseOperator(H)
    res = lanczos_ground_state(op, max_iters=400, tol=1e-12, compute_eigenvector=True, check_every=5)
    assert res.converged and abs(res.energy - w[0]) < 1e-10 * abs(w[0])
    vec = res.eigenvector.numpy()
    assert abs(abs(vec @ v[:, 0]) - 1) < 1e-8
    assert res.matvecs == 2 * res.iterations == op.products


def test_energy_tolerance_stops_before_the_residual_is_small():
    H = _matrix(600, 2)
    w = np.linalg.eigvalsh(H)
    strict = lanczos_ground_state(DenseOperator(H), max_iters=600, tol=1e-13, check_every=5)
    loose = lanczos_ground_state(DenseOperator(H), max_iters=600, tol=1e-13, energy_tol=1e-9, check_every=5)
    assert loose.converged and loose.iterations < strict.iterations
    assert abs(loose.energy - w[0]) < 1e-7 * abs(w[0]) and loose.residual > strict.residual


def test_time_limit_returns_the_current_ritz_value():
    H = _matrix(300, 3)
    res = lanczos_ground_state(DenseOperator(H), max_iters=300, tol=1e-14, check_every=5, time_limit_s=0.0)
    assert res.iterations == 5 and not res.converged and np.isfinite(res.energy)


def test_checkpoint_and_resume_are_bit_identical(tmp_path):
    H = _matrix(500, 4)
    full = lanczos_ground_state(DenseOperator(H), max_iters=120, tol=0.0, check_every=10)
    prefix = str(tmp_path / "run")
    first = lanczos_ground_state(DenseOperator(H), max_iters=50, tol=0.0, check_every=10, checkpoint=prefix,
                                 checkpoint_every=20)
    assert first.iterations == 50          # "the job ends here"; the last complete checkpoint is iteration 40
    op = DenseOperator(H)
    rest = lanczos_ground_state(op, max_iters=120, tol=0.0, check_every=10, checkpoint=prefix, checkpoint_every=20,
                                resume=True)
    assert op.products == 120 - 40
    assert rest.iterations == 120
    assert np.array_equal(rest.alphas, full.alphas) and np.array_equal(rest.betas, full.betas)
    assert rest.energy == full.energy
    with pytest.raises(ValueError, match="different run"):
        lanczos_ground_state(DenseOperator(H), max_iters=60, seed=7, checkpoint=prefix, resume=True)


@pytest.mark.parametrize("cplx", [False, True])
def test_thick_restart_lowest_pairs(cplx):
    H = _matrix(300, 5, cplx)
    w = np.linalg.eigvalsh(H)
    res = lanczos_thick_restart(DenseOperator(H), k=4, basis_size=40, tol=1e-11,
                                dtype=torch.complex128 if cplx else torch.float64)
    assert res.converged and np.allclose(res.energies, w[:4], rtol=0, atol=1e-9 * abs(w[0]))
    Ht = torch.from_numpy(H)
    for i in range(4):
        v = res.eigenvectors[i]
        assert float(torch.linalg.vector_norm(Ht.to(v.dtype) @ v - res.energies[i] * v)) < 1e-8 * abs(w[0])


# ---- block thick-restart Lanczos -----------------------------------------------------------------------------------------
def _check_pairs(H, res, k, atol=1e-9):
    w = np.linalg.eigvalsh(H)
    assert res.converged and np.allclose(res.energies, w[:k], rtol=0, atol=atol * max(1.0, abs(w[0])))
    Ht = torch.from_numpy(H)
    for i in range(k):
        v = res.eigenvectors[i]
        assert float(torch.linalg.vector_norm(Ht.to(v.dtype) @ v - res.energies[i] * v)) < 1e-7 * max(1.0, abs(w[0]))
    G = (res.eigenvectors.conj() @ res.eigenvectors.t()).numpy()
    assert np.allclose(G, np.eye(k), atol=1e-8)


@pytest.mark.parametrize("block", [1, 2, 4, 8])
def test_block_thick_restart_lowest_pairs(block):
    H = _matrix(400, 6)
    op = BlockDenseOperator(H)
    res = lanczos_block_thick_restart(op, k=4, block_size=block, basis_size=48, tol=1e-11)
    _check_pairs(H, res, 4)
    if block > 1:   # every step is ONE block product (a shrunken block of one vector goes through the single product)
        assert op.block_products > 0 and op.products + op.block_products < res.matvecs
    else:
        assert op.block_products == 0 and op.products == res.matvecs


def test_block_size_one_is_the_single_vector_method():
    H = _matrix(300, 7)
    single = lanczos_thick_restart(DenseOperator(H), k=3, basis_size=30, tol=1e-11)
    block = lanczos_block_thick_restart(DenseOperator(H), k=3, block_size=1, basis_size=30, tol=1e-11)
    assert np.allclose(single.energies, block.energies, atol=1e-10)


def test_block_method_resolves_a_degenerate_level():
    """Lowest level three-fold degenerate: a single vector sees one copy per Krylov space, a block of >= 3 sees all."""
    rng = np.random.default_rng(8)
    q, _ = np.linalg.qr(rng.standard_normal((200, 200)))
    levels = np.concatenate([[-5.0, -5.0, -5.0, -4.0, -3.5], np.linspace(-3, 3, 195)])
    H = (q * levels) @ q.T
    H = (H + H.T) / 2
    res = lanczos_block_thick_restart(BlockDenseOperator(H), k=5, block_size=4, basis_size=40, tol=1e-11)
    assert res.converged and np.allclose(res.energies, levels[:5], atol=1e-9)
    _check_pairs(H, res, 5)


def test_block_thick_restart_complex_falls_back_to_single_products():
    H = _matrix(200, 9, cplx=True)
    op = BlockDenseOperator(H)
    res = lanczos_block_thick_restart(op, k=3, block_size=3, basis_size=30, tol=1e-11, dtype=torch.complex128)
    _check_pairs(H, res, 3)
    assert op.block_products == 0 and op.products == res.matvecs


@pytest.mark.parametrize("dim,k,block", [(1, 1, 4), (3, 2, 4), (10, 10, 3), (24, 3, 8), (40, 5, 4)])
def test_block_thick_restart_small_spaces(dim, k, block):
    """The basis may fill the whole space: the residual block loses rank, the iteration stops with exact pairs."""
    H = _matrix(dim, 10 + dim)
    res = lanczos_block_thick_restart(BlockDenseOperator(H), k=k, block_size=block, tol=1e-12)
    _check_pairs(H, res, min(k, dim), atol=1e-10)


def test_block_thick_restart_with_many_restarts():
    H = _matrix(500, 11)
    res = lanczos_block_thick_restart(BlockDenseOperator(H), k=2, block_size=3, basis_size=14, tol=1e-10, max_restarts=2000)
    _check_pairs(H, res, 2)


# ---- several ranks (gloo): the solvers on row blocks, collectives supplied by the stand-in --------------------------------
class ShardedDenseOperator:
    """Rank ``rank`` of ``world`` holds a contiguous block of rows of a dense symmetric matrix; the product all-gathers
    x (the library's all-gather form), dots and Gram matrices are all-reduced -- what DistributedOperator does over NCCL."""
    device = "cpu"

    def __init__(self, matrix, rank, world, bounds=None):
        import torch.distributed as dist
        self.dist = dist
        dim = matrix.shape[0]
        bounds = bounds or [dim * r // world for r in range(world + 1)]
        self.bounds = bounds
        self.rows = torch.from_numpy(np.ascontiguousarray(matrix[bounds[rank]:bounds[rank + 1]]))
        self.layout = Layout(world, rank, dim, bounds[rank], bounds[rank + 1], 0, 0, 0, list(bounds))
        self.products = 0

    def empty_vector(self, dtype=None):
        return torch.zeros(self.layout.rows, dtype=dtype or torch.float64)

    def _gather(self, x):
        """All blocks of a vector (rows = last axis), on every rank: gloo's all_gather wants equal sizes, so every rank
        adds its block into a zero vector of the full length."""
        L = self.layout
        full = torch.zeros(x.shape[:-1] + (L.dim,), dtype=x.dtype)
        full[..., L.row_begin:L.row_end] = x
        self.dist.all_reduce(full)
        return full

    def matvec(self, x, y, mode=None):
        self.products += 1
        y.copy_(self.rows.to(x.dtype) @ self._gather(x))

    def matvec_block(self, X, Y):
        for x, y in zip(X, Y):
            self.matvec(x, y)

    def dot(self, a, b):
        return self.allreduce_(torch.vdot(a, b).reshape(1))

    def allreduce_(self, t):
        t = t.contiguous()
        self.dist.all_reduce(t)
        return t

    def sync(self):
        pass


def _worker_solvers(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H = _matrix(240, 21)
        w = np.linalg.eigvalsh(H)
        # uneven row ranges, one of them empty when there are three ranks (a re-balanced layout may look like that)
        bounds = [0, 100, 240] if world == 2 else [0, 100, 100, 240]
        ok = True
        op = ShardedDenseOperator(H, rank, world, bounds)
        plain = lanczos_ground_state(op, max_iters=300, tol=1e-12, check_every=5, compute_eigenvector=True)
        ok &= bool(plain.converged and abs(plain.energy - w[0]) < 1e-10 * abs(w[0]))
        thick = lanczos_thick_restart(ShardedDenseOperator(H, rank, world, bounds), k=3, basis_size=30, tol=1e-11)
        ok &= bool(thick.converged and np.allclose(thick.energies, w[:3], atol=1e-9 * abs(w[0])))
        block = lanczos_block_thick_restart(ShardedDenseOperator(H, rank, world, bounds), k=3, block_size=3, basis_size=30,
                                            tol=1e-11)
        ok &= bool(block.converged and np.allclose(block.energies, w[:3], atol=1e-9 * abs(w[0])))
        # the eigenvector blocks of all ranks, put together, are eigenvectors of the whole matrix
        for res in (thick, block):
            full = op._gather(res.eigenvectors).numpy()
            for e, v in zip(res.energies, full):
                ok &= bool(np.linalg.norm(H @ v - e * v) < 1e-8 * abs(w[0]))
        # the same numbers as a single rank finds: the start vector is a function of the GLOBAL row only
        alone = lanczos_thick_restart(DenseOperator(H), k=3, basis_size=30, tol=1e-11)
        ok &= bool(np.allclose(alone.energies, thick.energies, atol=1e-10) and alone.matvecs == thick.matvecs)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_solvers_on_row_blocks_under_gloo(world):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        out = manager.dict()
        procs = [ctx.Process(target=_worker_solvers, args=(r, world, port, out)) for r in range(world)]
        for pr in procs:
            pr.start()
        for pr in procs:
            pr.join(timeout=300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        assert dict(out) == {r: True for r in range(world)}
