"""Shared test plumbing: reference fixtures (hard-coded from the reference's own
test inputs, file:line cited) and the glue that feeds one problem description
to both the oracle and the CUDA path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from lattice_symmetries_b200.expr import Expr, compile_terms
from lattice_symmetries_b200.symmetry import Symmetries, Symmetry

HEIS_S = "Sˣ₀ Sˣ₁ + Sʸ₀ Sʸ₁ + Sᶻ₀ Sᶻ₁"
HOP = "- (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓)"


@dataclass
class Problem:
    name: str
    number_sites: int
    expr: Expr
    particle: int = 0                       # 0 spin, 1 spinful fermion, 2 spinless fermion
    hamming_weight: Optional[int] = None    # spin
    number_particles: Optional[object] = None   # fermions: None | int | (up, down)
    spin_inversion: Optional[int] = None
    symmetries: Optional[Symmetries] = None
    energy: Optional[float] = None          # known ground-state energy

    # -- oracle side ---------------------------------------------------------------
    def oracle_basis(self, oracle):
        if self.particle == 0:
            syms = self.symmetries if self.symmetries is not None else Symmetries([])
            group = oracle.Group.from_symmetries(syms, self.number_sites, self.spin_inversion)
            return oracle.Basis(self.number_sites, 0, self.number_sites, self.hamming_weight,
                                self.spin_inversion, group)
        total, up = None, None
        if isinstance(self.number_particles, (tuple, list)):
            up = int(self.number_particles[0])
            total = up + int(self.number_particles[1])
        elif self.number_particles is not None:
            total = int(self.number_particles)
        return oracle.Basis(self.number_sites, self.particle, total, up, None, None)

    def terms(self, oracle):
        ts = compile_terms(self.expr, self.number_sites)
        off = oracle.Terms([t for t in ts if t.x != 0])
        diag = oracle.Terms([t for t in ts if t.x == 0])
        return off, diag

    def oracle_setup(self, oracle):
        b = self.oracle_basis(oracle)
        reps = b.enumerate()
        index = oracle.Index(reps, b.number_bits, 22)
        off, diag = self.terms(oracle)
        return b, reps, index, off, diag

    # -- product side ----------------------------------------------------------------
    def product_basis(self):
        import lattice_symmetries_b200 as ls
        if self.particle == 0:
            return ls.SpinBasis(self.number_sites, self.hamming_weight, self.spin_inversion, self.symmetries)
        if self.particle == 1:
            return ls.SpinfulFermionBasis(self.number_sites, self.number_particles)
        return ls.SpinlessFermionBasis(self.number_sites, self.number_particles)


def oracle_ground_state_energy(oracle, problem: Problem) -> Tuple[float, int]:
    """Lowest eigenvalue of the oracle's matvec (dense for tiny dims, eigsh otherwise)."""
    import scipy.sparse.linalg as sla

    b, reps, index, off, diag = problem.oracle_setup(oracle)
    dim = reps.shape[0]

    def mv(x):
        return oracle.matvec(b, off, diag, index, np.ascontiguousarray(x, dtype=np.float64).reshape(-1))[0]

    if dim <= 600:
        H = np.stack([mv(e) for e in np.eye(dim)], axis=1)
        assert np.allclose(H, H.T, atol=1e-12)
        return float(np.linalg.eigvalsh(H)[0]), dim
    op = sla.LinearOperator((dim, dim), matvec=mv, dtype=np.float64)
    w = sla.eigsh(op, k=1, which="SA", tol=1e-10)[0]
    return float(w[0]), dim


# ---- the reference's own known answers ---------------------------------------------------
def chain10_getting_started() -> Problem:
    """python/example/getting_started.py:12-51"""
    n = 10
    T = Symmetry([(i + 1) % n for i in range(n)], sector=n // 2)
    P = Symmetry(list(range(n))[::-1], sector=1)
    edges = [(i, (i + 1) % n) for i in range(n)]
    return Problem("chain10", n, Expr("2 (σ⁺₀ σ⁻₁ + σ⁺₁ σ⁻₀) + σᶻ₀ σᶻ₁", sites=edges), hamming_weight=n // 2,
                   spin_inversion=-1, symmetries=Symmetries([T, P]), energy=-18.06178542)


def hphi_01_kagome() -> Problem:
    """test/01_spin_kagome/hamiltonian.yaml; HPhi/output/zvo_energy.dat:1"""
    e = (Expr(HEIS_S, sites=[[1, 2], [1, 5], [4, 5], [4, 8], [7, 8], [7, 2]])
         + Expr("0.5 (" + HEIS_S + ")", sites=[[0, 2], [2, 3], [3, 5], [5, 6], [6, 8], [8, 0]])
         + Expr("0.5 (" + HEIS_S + ")", sites=[[0, 1], [1, 6], [3, 4], [4, 0], [6, 7], [7, 3]]))
    return Problem("hphi01", 9, e, hamming_weight=5, energy=-3.0170209179017471)


def hphi_02_ladder() -> Problem:
    """test/02_spin_ladder_DM/hamiltonian.yaml"""
    e = (Expr(HEIS_S, sites=[[0, 2], [0, 4], [0, 1], [1, 3], [1, 5], [2, 4], [2, 3], [3, 5], [4, 5]])
         + Expr("0.5 Sˣ₀", sites=[[i] for i in range(6)]))
    return Problem("hphi02", 6, e, energy=-3.0527756377319952)


def hphi_03_hcor() -> Problem:
    """test/03_spin_hcor/hamiltonian.yaml"""
    e = (Expr("- Sˣ₀ Sˣ₁", sites=[[0, 5], [1, 4], [2, 7], [3, 6]])
         + Expr("- Sʸ₀ Sʸ₁", sites=[[0, 3], [1, 2], [4, 7], [5, 6]])
         + Expr("- Sᶻ₀ Sᶻ₁", sites=[[0, 1], [2, 3], [4, 5], [6, 7]])
         + Expr("- Sᶻ₀", sites=[[i] for i in range(8)]))
    return Problem("hphi03", 8, e, energy=-5.1653788071251920)


def hphi_04_hubbard_square() -> Problem:
    """test/04_hubbard_square/hamiltonian.yaml"""
    bonds = [[0, 1], [0, 3], [0, 4], [1, 2], [1, 5], [2, 3], [2, 6], [3, 7], [4, 5], [4, 7], [5, 6], [6, 7],
             [0, 4], [1, 5], [2, 6], [3, 7]]
    e = Expr(HOP, sites=bonds) + Expr("4.0 n₀↑ n₀↓", sites=[[i] for i in range(8)])
    return Problem("hphi04", 8, e, particle=1, number_particles=(4, 4), energy=-10.2529529552635879)


def hphi_05_hubbard_tri() -> Problem:
    """test/05_hubbard_tri/hamiltonian.yaml"""
    e = (Expr(HOP, sites=[[3, 0], [0, 3], [4, 1], [1, 4], [5, 2], [2, 5], [1, 3], [3, 1], [2, 4], [4, 2], [0, 5],
                          [5, 0], [0, 1], [1, 2], [0, 2], [3, 4], [4, 5], [3, 5]])
         + Expr("- 2 (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓)", sites=[[0, 4], [1, 5], [2, 3]])
         + Expr("- (c†₀↑ c₀↑ + c†₀↓ c₀↓)", sites=[[i] for i in range(6)])
         + Expr("4.0 n₀↑ n₀↓", sites=[[i] for i in range(6)]))
    return Problem("hphi05", 6, e, particle=1, number_particles=None, energy=-17.4356927965492972)


KAGOME12_EXPR = (
    "1.0 σᶻ₀ σᶻ₁ + 1.0 σᶻ₀ σᶻ₃ + 1.0 σᶻ₀ σᶻ₈ + 1.0 σᶻ₀ σᶻ₁₀ + 2.0 σ⁺₀ σ⁻₁ + 2.0 σ⁺₀ σ⁻₃ + 2.0 σ⁺₀ σ⁻₈ + 2.0 σ⁺₀ σ⁻₁₀ + 2.0 σ⁻₀ σ⁺₁ + 2.0 σ⁻₀ σ⁺₃ + 2.0 σ⁻₀ σ⁺₈ + 2.0 σ⁻₀ σ⁺₁₀ + 1.0 σᶻ₁ σᶻ₂ + 0.8 σᶻ₁ σᶻ₃ + 0.8 σᶻ₁ σᶻ₉ + 2.0 σ⁺₁ σ⁻₂ + 1.6 σ⁺₁ σ⁻₃ + 1.6 σ⁺₁ σ⁻₉ + 2.0 σ⁻₁ σ⁺₂ + 1.6 σ⁻₁ σ⁺₃ + 1.6 σ⁻₁ σ⁺₉ + 1.0 σᶻ₂ σᶻ₄ + 1.0 σᶻ₂ σᶻ₉ + 1.0 σᶻ₂ σᶻ₁₀ + 2.0 σ⁺₂ σ⁻₄ + 2.0 σ⁺₂ σ⁻₉ + 2.0 σ⁺₂ σ⁻₁₀ + 2.0 σ⁻₂ σ⁺₄ + 2.0 σ⁻₂ σ⁺₉ + 2.0 σ⁻₂ σ⁺₁₀ + 1.0 σᶻ₃ σᶻ₅ + 0.8 σᶻ₃ σᶻ₁₁ + 2.0 σ⁺₃ σ⁻₅ + 1.6 σ⁺₃ σ⁻₁₁ + 2.0 σ⁻₃ σ⁺₅ + 1.6 σ⁻₃ σ⁺₁₁ + 0.8 σᶻ₄ σᶻ₆ + 1.0 σᶻ₄ σᶻ₇ + 0.8 σᶻ₄ σᶻ₁₀ + 1.6 σ⁺₄ σ⁻₆ + 2.0 σ⁺₄ σ⁻₇ + 1.6 σ⁺₄ σ⁻₁₀ + 1.6 σ⁻₄ σ⁺₆ + 2.0 σ⁻₄ σ⁺₇ + 1.6 σ⁻₄ σ⁺₁₀ + 1.0 σᶻ₅ σᶻ₆ + 1.0 σᶻ₅ σᶻ₈ + 1.0 σᶻ₅ σᶻ₁₁ + 2.0 σ⁺₅ σ⁻₆ + 2.0 σ⁺₅ σ⁻₈ + 2.0 σ⁺₅ σ⁻₁₁ + 2.0 σ⁻₅ σ⁺₆ + 2.0 σ⁻₅ σ⁺₈ + 2.0 σ⁻₅ σ⁺₁₁ + 1.0 σᶻ₆ σᶻ₇ + 0.8 σᶻ₆ σᶻ₈ + 2.0 σ⁺₆ σ⁻₇ + 1.6 σ⁺₆ σ⁻₈ + 2.0 σ⁻₆ σ⁺₇ + 1.6 σ⁻₆ σ⁺₈ + 1.0 σᶻ₇ σᶻ₉ + 1.0 σᶻ₇ σᶻ₁₁ + 2.0 σ⁺₇ σ⁻₉ + 2.0 σ⁺₇ σ⁻₁₁ + 2.0 σ⁻₇ σ⁺₉ + 2.0 σ⁻₇ σ⁺₁₁ + 0.8 σᶻ₈ σᶻ₁₀ + 1.6 σ⁺₈ σ⁻₁₀ + 1.6 σ⁻₈ σ⁺₁₀ + 0.8 σᶻ₉ σᶻ₁₁ + 1.6 σ⁺₉ σ⁻₁₁ + 1.6 σ⁻₉ σ⁺₁₁"
)


def kagome12_complex_sector() -> Problem:
    """python/test/test_api.py:45-61: 12-site kagome with one translation in the
    complex sector 1 (the reference asserts only that eigsh runs)."""
    right_shift = Symmetry([2, 10, 0, 4, 3, 7, 11, 5, 9, 8, 1, 6], sector=1)
    return Problem("kagome12", 12, Expr(KAGOME12_EXPR), hamming_weight=6, symmetries=Symmetries([right_shift]))


def hubbard2(number_particles=None, t=1, U=2) -> Problem:
    """python/run_tests.py:79-88 create_hubbard_hamiltonian"""
    op = -t * Expr("c†↑₀ c↑₁", [(0, 1)])
    op = op - t * Expr("c†↑₁ c↑₀", [(0, 1)])
    op = op - t * Expr("c†↓₁ c↓₀", [(0, 1)])
    op = op - t * Expr("c†↓₀ c↓₁", [(0, 1)])
    op = op + U * Expr("n↑₀ n↓₀", [(0,)])
    op = op + U * Expr("n↑₁ n↓₁", [(1,)])
    return Problem("hubbard2", 2, op, particle=1, number_particles=number_particles)


HUBBARD2_MATRIX_16 = np.zeros((16, 16))
for _i, _j, _v in [(1, 2, -1), (2, 1, -1), (4, 8, -1), (8, 4, -1), (5, 5, 2), (5, 6, -1), (5, 9, -1), (6, 5, -1),
                   (6, 10, -1), (7, 7, 2), (7, 11, -1), (9, 5, -1), (9, 10, -1), (10, 6, -1), (10, 9, -1),
                   (10, 10, 2), (11, 7, -1), (11, 11, 2), (13, 13, 2), (13, 14, -1), (14, 13, -1), (14, 14, 2),
                   (15, 15, 4)]:
    HUBBARD2_MATRIX_16[_i, _j] = _v  # python/run_tests.py:128-147

HUBBARD2_MATRIX_6 = np.array([  # python/run_tests.py:160-169
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 2.0, -1.0, -1.0, 0.0, 0.0],
    [0.0, -1.0, 0.0, 0.0, -1.0, 0.0],
    [0.0, -1.0, 0.0, 0.0, -1.0, 0.0],
    [0.0, 0.0, -1.0, -1.0, 2.0, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
])


def dense_from_rows(apply_diag, apply_off_diag, states) -> np.ndarray:
    """python/run_tests.py:91-118 calculate_hamiltonian_matrix, generic in the
    provider of the two row queries."""
    states = [int(s) for s in states]
    pos = {s: i for i, s in enumerate(states)}
    H = np.zeros((len(states), len(states)))
    for ket in states:
        c = apply_diag(ket)
        if c != 0:
            H[pos[ket], pos[ket]] += c
        for coeff, bra in apply_off_diag(ket):
            if coeff != 0 and bra in pos:
                assert complex(coeff).imag == 0
                H[pos[bra], pos[ket]] += complex(coeff).real
    return H


def random_fixed_hamming_states(rng, number_bits: int, hamming_weight: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint64)
    for i in range(count):
        bits = rng.choice(number_bits, size=hamming_weight, replace=False)
        out[i] = np.bitwise_or.reduce(np.uint64(1) << bits.astype(np.uint64))
    return out


# ---- model files and the stand-in operator of the exact-diagonalisation program -------------------------------------------
CHAIN10_YAML = """\
# python/example/getting_started.py:12-51 as a model file (E0 = -18.06178542, dim 13)
basis:
  number_spins: 10
  hamming_weight: 5
  spin_inversion: -1
  symmetries:
    - permutation: [1, 2, 3, 4, 5, 6, 7, 8, 9, 0]
      sector: 5
    - permutation: [9, 8, 7, 6, 5, 4, 3, 2, 1, 0]
      sector: 1
hamiltonian:
  name: "Heisenberg Hamiltonian"
  lattice: &lattice [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 8], [8, 9], [9, 0]]
  terms:
    - expression: "2 (σ⁺₀ σ⁻₁ + σ⁺₁ σ⁻₀)"
      sites: *lattice
    - expression: "σᶻ₀ σᶻ₁"
      sites: *lattice
observables:
  - terms:
      - expression: "σᶻ₀ σᶻ₁"
        sites: [[0, 5]]
number_vectors: 2
output: "chain10.h5"
"""


def problem_of(parsed) -> Problem:
    m = parsed.model
    particle = {"spin-1/2": 0, "spinful-fermion": 1, "spinless-fermion": 2}[m.particle]
    return Problem(m.name, m.number_sites, parsed.hamiltonian, particle=particle, hamming_weight=m.hamming_weight,
                   number_particles=m.number_particles, spin_inversion=m.spin_inversion, symmetries=m.symmetries)


class OracleOperator:
    """The solver interface (lanczos._wrap) over the oracle's CPU matvec."""
    device = "cpu"

    def __init__(self, oracle, problem):
        self.oracle = oracle
        self.b, self.reps, self.index, self.off, self.diag = problem.oracle_setup(oracle)
        dim = int(self.reps.shape[0])
        from lattice_symmetries_b200.distributed import Layout
        self.layout = Layout(1, 0, dim, 0, dim, 0, 0, 0, [0, dim])

    def empty_vector(self, dtype=None):
        import torch
        return torch.zeros(self.layout.dim, dtype=dtype or torch.float64)

    def matvec(self, x, y, mode=None):
        import torch
        out = self.oracle.matvec(self.b, self.off, self.diag, self.index, np.ascontiguousarray(x.numpy()))[0]
        y.copy_(torch.from_numpy(np.ascontiguousarray(out)))

    def matvec_block(self, X, Y):
        self.block_products = getattr(self, "block_products", 0) + 1
        for x, y in zip(X, Y):
            self.matvec(x, y)

    def dot(self, a, b):
        import torch
        return torch.vdot(a, b).reshape(1)

    def sync(self):
        pass

    def dense(self) -> np.ndarray:
        import torch
        dim = self.layout.dim
        out = torch.zeros(dim, dtype=torch.float64)
        cols = []
        for e in torch.eye(dim, dtype=torch.float64):
            self.matvec(e, out)
            cols.append(out.numpy().copy())
        return np.stack(cols, axis=1)




def yaml_of_model(model, extra: str = "") -> str:
    """A Model of lattices.py as the reference's YAML (spin bases with bonds)."""
    lines = ["basis:", f"  number_spins: {model.number_sites}"]
    if model.hamming_weight is not None:
        lines.append(f"  hamming_weight: {model.hamming_weight}")
    if model.spin_inversion is not None:
        lines.append(f"  spin_inversion: {model.spin_inversion}")
    if model.symmetries is not None and len(model.symmetries):
        lines.append("  symmetries:")
        for g in model.symmetries.generators:
            lines.append(f"    - permutation: {[int(i) for i in g.permutation]}")
            lines.append(f"      sector: {g.sector}")
    bonds = [[int(a), int(b)] for a, b in model.bonds]
    lines += ["hamiltonian:", '  name: "Heisenberg Hamiltonian"', f"  lattice: &lattice {bonds}", "  terms:"]
    for e in ("σˣ₀ σˣ₁", "σʸ₀ σʸ₁", "σᶻ₀ σᶻ₁"):
        lines += [f'    - expression: "{e}"', "      sites: *lattice"]
    return "\n".join(lines) + "\n" + extra
