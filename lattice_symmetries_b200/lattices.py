"""Synthetic models of the shapes named in BASELINE.json ``configs``:
lattices, symmetry generators and Hamiltonian expressions.

The reference ships its models as YAML (chapel/data/*.yaml, test/0N_*/hamiltonian.yaml);
the large shapes (kagome 36 / 42, 2x16 ladder with DM, 4x4 Hubbard) are not in
its tree and are synthesised here.  Everything in this module is host-side
set-up; it only *describes* inputs for the CUDA path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .expr import Expr
from .symmetry import Symmetries, Symmetry

__all__ = [
    "Model", "heisenberg_chain", "kagome_cluster", "kagome_heisenberg", "ladder_dm", "hubbard_square",
    "chain_symmetries", "HEISENBERG_BOND",
]

HEISENBERG_BOND = "σˣ₀ σˣ₁ + σʸ₀ σʸ₁ + σᶻ₀ σᶻ₁"


@dataclass
class Model:
    """Inputs of one benchmark configuration."""
    name: str
    number_sites: int
    expression: Expr
    hamming_weight: Optional[int] = None
    spin_inversion: Optional[int] = None
    symmetries: Optional[Symmetries] = None
    particle: str = "spin-1/2"
    number_particles: Optional[object] = None   # fermions: None | N | (N_up, N_down)
    bonds: List[Tuple[int, int]] = field(default_factory=list)

    def basis(self):
        from .basis import SpinBasis, SpinfulFermionBasis, SpinlessFermionBasis
        if self.particle == "spin-1/2":
            return SpinBasis(self.number_sites, self.hamming_weight, self.spin_inversion, self.symmetries)
        if self.particle == "spinless-fermion":
            return SpinlessFermionBasis(self.number_sites, self.number_particles)
        return SpinfulFermionBasis(self.number_sites, self.number_particles)

    def operator(self, basis=None):
        from .operator import Operator
        return Operator(basis if basis is not None else self.basis(), self.expression)


# ---------------------------------------------------------------------------------
def chain_symmetries(n: int, translation_sector: Optional[int] = 0, parity_sector: Optional[int] = 0) -> Symmetries:
    gens = []
    if translation_sector is not None:
        gens.append(Symmetry([(i + 1) % n for i in range(n)], translation_sector))
    if parity_sector is not None:
        gens.append(Symmetry(list(range(n))[::-1], parity_sector))
    return Symmetries(gens)


def heisenberg_chain(n: int, symmetric: bool = True, translation_sector: int = 0, parity_sector: int = 0,
                     spin_inversion: Optional[int] = 1) -> Model:
    """chapel/data/heisenberg_chain_{n}_symm.yaml: ring of n spins, Sz = 0,
    T = [1..n-1, 0], P = reverse, spin inversion."""
    bonds = [(i, (i + 1) % n) for i in range(n)]
    expr = Expr(HEISENBERG_BOND, sites=bonds)
    if not symmetric:
        return Model(f"heisenberg_chain_{n}", n, expr, hamming_weight=n // 2, bonds=bonds)
    return Model(f"heisenberg_chain_{n}_symm", n, expr, hamming_weight=n // 2, spin_inversion=spin_inversion,
                 symmetries=chain_symmetries(n, translation_sector, parity_sector), bonds=bonds)


# ---------------------------------------------------------------------------------
class _Cluster:
    """Sites of a 2-D lattice with a basis, folded into the torus spanned by
    the integer superlattice vectors T1, T2 (given in units of a1, a2).
    Positions are kept in exact half-integer lattice coordinates (doubled)."""

    def __init__(self, T1: Tuple[int, int], T2: Tuple[int, int], basis2: Sequence[Tuple[int, int]]):
        self.T = np.array([T1, T2], dtype=np.int64)            # rows: superlattice vectors
        self.det = int(round(abs(np.linalg.det(self.T))))
        self.basis2 = [tuple(b) for b in basis2]                 # doubled coordinates of the basis sites
        # enumerate cells: lattice points modulo the superlattice
        cells = []
        seen = set()
        span = abs(self.T).sum() + 1
        for i in range(-span, span + 1):
            for j in range(-span, span + 1):
                key = self._fold_cell((i, j))
                if key not in seen:
                    seen.add(key)
                    cells.append(key)
        cells.sort()
        assert len(cells) == self.det, (len(cells), self.det)
        self.cells = cells
        self.index = {}
        for c, cell in enumerate(cells):
            for b in range(len(self.basis2)):
                self.index[(cell, b)] = c * len(self.basis2) + b
        self.number_sites = len(cells) * len(self.basis2)

    def _fold_cell(self, v: Tuple[int, int]) -> Tuple[int, int]:
        # v = c1 T1 + c2 T2 with rational c; reduce c mod 1 exactly with integers
        T = self.T
        d = T[0, 0] * T[1, 1] - T[0, 1] * T[1, 0]
        c1 = v[0] * T[1, 1] - v[1] * T[1, 0]   # times d
        c2 = -v[0] * T[0, 1] + v[1] * T[0, 0]  # times d
        n1 = int(np.floor(c1 / d)) if d > 0 else int(np.floor(-c1 / -d))
        n2 = int(np.floor(c2 / d)) if d > 0 else int(np.floor(-c2 / -d))
        return (int(v[0] - n1 * T[0, 0] - n2 * T[1, 0]), int(v[1] - n1 * T[0, 1] - n2 * T[1, 1]))

    def site(self, pos2: Tuple[int, int]) -> int:
        """Site index of the doubled lattice coordinate pos2 (folded)."""
        for b, off in enumerate(self.basis2):
            d0, d1 = pos2[0] - off[0], pos2[1] - off[1]
            if d0 % 2 == 0 and d1 % 2 == 0:
                return self.index[(self._fold_cell((d0 // 2, d1 // 2)), b)]
        raise ValueError(f"{pos2} is not a lattice site")

    def positions2(self) -> List[Tuple[int, int]]:
        out = [None] * self.number_sites
        for (cell, b), i in self.index.items():
            out[i] = (2 * cell[0] + self.basis2[b][0], 2 * cell[1] + self.basis2[b][1])
        return out

    def permutation(self, f: Callable[[Tuple[int, int]], Tuple[int, int]]) -> List[int]:
        """Permutation p with p[i] = index of the image of site i under the
        lattice map f (doubled coordinates)."""
        pos = self.positions2()
        p = [self.site(f(r)) for r in pos]
        assert sorted(p) == list(range(self.number_sites)), "map is not a symmetry of the cluster"
        return p


def kagome_cluster(T1: Tuple[int, int], T2: Tuple[int, int]):
    """Kagome lattice: triangular Bravais lattice (a1, a2 at 60 degrees) with
    sites A = 0, B = a1/2, C = a2/2.  Returns (cluster, bonds)."""
    cl = _Cluster(T1, T2, [(0, 0), (1, 0), (0, 1)])
    bonds = set()
    for r in cl.positions2():
        if r[0] % 2 == 0 and r[1] % 2 == 0:  # an A site: its up and down triangles
            up = [r, (r[0] + 1, r[1]), (r[0], r[1] + 1)]
            dn = [r, (r[0] - 1, r[1]), (r[0], r[1] - 1)]
            for tri in (up, dn):
                s = [cl.site(q) for q in tri]
                for a in range(3):
                    for b in range(a + 1, 3):
                        bonds.add((min(s[a], s[b]), max(s[a], s[b])))
    return cl, sorted(bonds)


def _kagome_generators(cl: _Cluster, point_group: str):
    """Symmetry generators as lattice maps in doubled coordinates (n1, n2)
    of r = (n1 a1 + n2 a2) / 2.  The C6 centre is the hexagon centre (1, 1)."""
    gens = [lambda r: (r[0] + 2, r[1]), lambda r: (r[0], r[1] + 2)]  # translations by a1, a2

    def about_centre(m):
        def f(r):
            x, y = r[0] - 1, r[1] - 1
            u, v = m(x, y)
            return (u + 1, v + 1)
        return f

    if point_group in ("c6v", "c6"):
        gens.append(about_centre(lambda x, y: (-y, x + y)))      # rotation by 60 degrees: a1 -> a2, a2 -> a2 - a1
    if point_group in ("c2", "c2v"):
        gens.append(about_centre(lambda x, y: (-x, -y)))          # rotation by 180 degrees
    if point_group in ("c6v", "c2v"):
        gens.append(about_centre(lambda x, y: (y, x)))            # mirror exchanging a1 and a2
    return [cl.permutation(g) for g in gens]


def kagome_heisenberg(number_sites: int = 36, spin_inversion: Optional[int] = None,
                      point_group: Optional[str] = None) -> Model:
    """Kagome Heisenberg antiferromagnet, Sz = 0, all symmetry sectors trivial.

    36 sites: 12 cells spanned by T1 = 2 a1 + 2 a2, T2 = -2 a1 + 4 a2 (C6v cluster,
    |G| = 12 x 12 = 144).  42 sites: 14 cells spanned by T1 = 4 a1 + a2... (14 is not a
    Loeschian number, so there is no C6 cluster; C2 about the hexagon centre always
    survives).  12 / 18 / 24 / 27 / 30-site clusters are provided for tests."""
    shapes = {
        12: ((2, 0), (0, 2), "c6v"),
        18: ((3, 0), (0, 2), "c2"),
        24: ((2, 2), (-2, 2), "c2v"),
        27: ((3, 0), (0, 3), "c6v"),
        30: ((5, 0), (0, 2), "c2"),
        36: ((2, 2), (-2, 4), "c6v"),
        42: ((7, 0), (2, 2), "c2"),
        48: ((4, 0), (0, 4), "c6v"),
    }
    T1, T2, default_pg = shapes[number_sites]
    pg = point_group if point_group is not None else default_pg
    cl, bonds = kagome_cluster(T1, T2)
    assert cl.number_sites == number_sites and len(bonds) == 2 * number_sites
    perms = _kagome_generators(cl, pg)
    for p in perms:  # every generator must map bonds to bonds
        image = {(min(p[a], p[b]), max(p[a], p[b])) for a, b in bonds}
        assert image == set(bonds), "generator does not preserve the bonds"
    # Symmetry(perm): output bit i = input bit perm[i]; the inverse map is an
    # equally good generator of the same group.
    syms = Symmetries([Symmetry(p, 0) for p in perms])
    expr = Expr(HEISENBERG_BOND, sites=bonds)
    return Model(f"kagome_{number_sites}_heisenberg", number_sites, expr, hamming_weight=number_sites // 2,
                 spin_inversion=spin_inversion, symmetries=syms, bonds=bonds)


# ---------------------------------------------------------------------------------
def ladder_dm(length: int = 16, D: float = 0.3, sector: int = 1) -> Model:
    """2 x length spin ladder, sites i = 2 x + leg, Heisenberg J = 1 on legs and
    rungs plus a Dzyaloshinskii-Moriya term D z.(S_i x S_j) = (i D / 2)(S+_i S-_j - S-_i S+_j)
    on the leg bonds, periodic along x; symmetry = translation by one rung in
    momentum sector ``sector`` (complex characters)."""
    n = 2 * length
    legs = [(2 * x + leg, 2 * ((x + 1) % length) + leg) for x in range(length) for leg in range(2)]
    rungs = [(2 * x, 2 * x + 1) for x in range(length)]
    heis = Expr("Sˣ₀ Sˣ₁ + Sʸ₀ Sʸ₁ + Sᶻ₀ Sᶻ₁", sites=legs + rungs)
    # S+ = sigma+ here (the reference's "S" prefix would also halve S+, Expr.hs:332-333)
    dm = Expr("σ⁺₀ σ⁻₁ - σ⁻₀ σ⁺₁", sites=legs).scale(0.5j * D)
    t = [0] * n
    for x in range(length):
        for leg in range(2):
            t[2 * x + leg] = 2 * ((x + 1) % length) + leg
    syms = Symmetries([Symmetry(t, sector)])
    return Model(f"ladder_2x{length}_dm", n, heis + dm, hamming_weight=n // 2, symmetries=syms, bonds=legs + rungs)


def hubbard_square(lx: int = 4, ly: int = 4, t: float = 1.0, U: float = 4.0,
                   number_particles: Optional[Tuple[int, int]] = None) -> Model:
    """lx x ly periodic square-lattice Hubbard model, expression as in
    test/04_hubbard_square/hamiltonian.yaml:7-13."""
    n = lx * ly
    bonds = set()
    for x in range(lx):
        for y in range(ly):
            i = x * ly + y
            for j in (((x + 1) % lx) * ly + y, x * ly + (y + 1) % ly):
                if i != j:
                    bonds.add((min(i, j), max(i, j)))
    bonds = sorted(bonds)
    hop = Expr("c†₀↑ c₁↑ + c†₁↑ c₀↑ + c†₀↓ c₁↓ + c†₁↓ c₀↓", sites=bonds).scale(-t)
    inter = Expr("n₀↑ n₀↓", sites=[[i] for i in range(n)]).scale(U)
    if number_particles is None:
        number_particles = (n // 2, n // 2)
    return Model(f"hubbard_{lx}x{ly}", n, hop + inter, particle="spinful-fermion",
                 number_particles=tuple(number_particles), bonds=bonds)
