"""Ground state of the 10-site Heisenberg ring in the (k = pi, parity-odd, spin-inversion-odd) sector on a B200.

The flow of the reference's python/example/getting_started.py with this package in place of ``lattice_symmetries``:
symmetries -> group -> symmetry-adapted basis (built on the GPU) -> expression -> operator -> ``eigsh`` driving the
library's matvec through the ``scipy.sparse.linalg.LinearOperator`` interface.  The energy is the one the reference
asserts (python/example/getting_started.py:51).  Like the reference's example it gives the basis NO Hamming weight: the
basis then holds every magnetisation sector compatible with the symmetries (34 states), and the ground state is found in
the S^z = 0 one.

    python examples/getting_started.py
"""
from __future__ import annotations

import sys
from functools import reduce
from pathlib import Path
import operator

import numpy as np
import scipy.sparse.linalg

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lattice_symmetries_b200 as ls  # noqa: E402

REFERENCE_ENERGY = -18.06178542


def main(verbose: bool = True) -> float:
    say = print if verbose else (lambda *a: None)
    number_spins = 10
    sites = np.arange(number_spins)
    translation = ls.Symmetry((sites + 1) % number_spins, sector=number_spins // 2)   # momentum pi
    parity = ls.Symmetry(sites[::-1], sector=1)                                       # odd under reflection
    symmetries = ls.Symmetries([translation, parity])
    say(f"{len(symmetries)} generators, {len(symmetries.elements)} group elements")

    basis = ls.SpinBasis(number_spins=number_spins, spin_inversion=-1, symmetries=symmetries)
    basis.build()
    say(f"Hilbert space dimension is {basis.number_states}")

    edges = [(i, (i + 1) % number_spins) for i in range(number_spins)]
    bond = "2 (σ⁺₀ σ⁻₁ + σ⁺₁ σ⁻₀) + σᶻ₀ σᶻ₁"
    expr = ls.Expr(bond, sites=edges)
    # the same expression, built algebraically
    summed = reduce(operator.add, (ls.Expr(bond).replace_indices({0: i, 1: j}) for i, j in edges))
    assert expr == summed

    hamiltonian = ls.Operator(basis, expr)
    eigenvalues, _ = scipy.sparse.linalg.eigsh(hamiltonian, k=1, which="SA")
    say(f"Ground state energy is {eigenvalues[0]}")
    assert np.isclose(eigenvalues[0], REFERENCE_ENERGY)
    return float(eigenvalues[0])


if __name__ == "__main__":
    main()
