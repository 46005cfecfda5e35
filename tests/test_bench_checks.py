"""CPU test of the full-size parity checker that every bench.py line carries (`checks.sampled_rows`): with a stand-in
basis (states + index answered by the oracle) and y computed column-wise from the oracle's primitives, the checker
must report rounding-level error -- and must SEE a perturbation of 1e-9, a wrong row, a missing representative."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _setup(model):
    import torch
    import bench
    from lattice_symmetries_b200.distributed import Layout, hashed_vector
    oracle, ob, off, diag = bench.oracle_problem(model)
    reps = ob.enumerate()
    index = oracle.Index(reps, ob.number_bits, 22)
    dim = reps.shape[0]

    class StandInBasis:
        states = reps

        def index(self, x):
            return index(x)

    complex_vectors = model.symmetries is not None and not bool(np.all(np.abs(model.symmetries.characters()[1]) < 1e-9))
    x = hashed_vector(0, dim, 42, device="cpu").numpy().astype(np.complex128)
    if complex_vectors:
        x = x + 1j * hashed_vector(0, dim, 42 + 1000003, device="cpu").numpy()
    betas, coeffs, offsets = oracle.apply_off_diag(off, reps)
    col = np.repeat(np.arange(dim), np.diff(offsets))
    if ob.c.has_permutation_symmetries:
        rb, rc, rn = ob.group.state_info(betas)
        na = ob.group.state_info(reps)[2]
        val, live = coeffs * rc * rn / na[col], rn > 0
    else:
        rb, val, live = betas, coeffs, np.ones(betas.shape[0], bool)
    y = np.zeros(dim, dtype=np.complex128)
    np.add.at(y, index(rb)[live], (val * x[col])[live])
    if diag.n:
        y += oracle.apply_diag(diag, reps) * x
    if not complex_vectors:
        y = y.real.copy()
    lay = Layout(1, 0, dim, 0, dim, 0, 0, 0, [0, dim])
    return bench, StandInBasis(), lay, torch.from_numpy(y), complex_vectors, index


@pytest.mark.parametrize("make", ["chain16", "kagome18", "ladder_dm8", "hubbard_2x3"])
def test_sampled_rows_checker_sees_errors(make):
    from lattice_symmetries_b200 import lattices as L
    model = {"chain16": lambda: L.heisenberg_chain(16), "kagome18": lambda: L.kagome_heisenberg(18),
             "ladder_dm8": lambda: L.ladder_dm(8), "hubbard_2x3": lambda: L.hubbard_square(2, 3)}[make]()
    bench, basis, lay, y, cplx, index = _setup(model)
    good = bench.sampled_rows_check(model, basis, lay, y, 42, cplx, 1, samples=512)
    assert good["missing"] == 0 and good["max_rel_err"] < 1e-13, good
    bad = bench.sampled_rows_check(model, basis, lay, y * (1 + 1e-9), 42, cplx, 1, samples=512)
    assert 1e-10 < bad["max_rel_err"] < 1e-7
    shifted = y.clone()
    shifted[1:] = y[:-1]
    assert bench.sampled_rows_check(model, basis, lay, shifted, 42, cplx, 1, samples=512)["max_rel_err"] > 1e-3

    class Lossy:   # an index that loses one representative in a hundred: reported as missing
        states = basis.states

        def index(self, x):
            j = index(x)
            j[::100] = -1
            return j
    assert bench.sampled_rows_check(model, Lossy(), lay, y, 42, cplx, 1, samples=512)["missing"] > 0
