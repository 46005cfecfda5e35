"""CPU tests of the Lanczos drivers on a stand-in operator: the recurrence, the stopping rules and the
checkpoint / resume logic are host code and do not care where the vectors live.  (On the GPU the same drivers run on
the library's products: tests/test_gpu_parity.py, tests/nccl_worker.py.)"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from lattice_symmetries_b200.distributed import Layout
from lattice_symmetries_b200.lanczos import lanczos_ground_state, lanczos_thick_restart


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
