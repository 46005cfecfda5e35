"""GPU tests of the callers either side of the hot path: a basis populated through ``ls_hs_unchecked_set_representatives``
(kernels/reference.c:196-211 -- what the reference's Chapel driver does on every locale, chapel/src/Diagonalize.chpl:293),
the YAML model reader feeding the library (python/lattice_symmetries/__init__.py:762-772) and the exact-diagonalisation
program (YAML in, HDF5 out: chapel/src/Diagonalize.chpl:258-333) on the library's products.  Checked against the CPU
oracle.  The model files are written by the tests themselves (/root/reference does not exist on the GPU box).
Run with ``pytest -m gpu`` on a B200."""
from __future__ import annotations

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


# ---- ls_hs_unchecked_set_representatives: borrowed representatives, norms computed on first use -----------------------
@pytest.mark.parametrize("make", [
    lambda L: L.heisenberg_chain(24), lambda L: L.kagome_heisenberg(18), lambda L: L.kagome_heisenberg(24, spin_inversion=1),
    lambda L: L.heisenberg_chain(20, translation_sector=3, parity_sector=None, spin_inversion=None)])
def test_unchecked_set_representatives_then_products(oracle, make):
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    model = make(L)
    p = H.Problem(model.name, model.number_sites, model.expression, hamming_weight=model.hamming_weight,
                  spin_inversion=model.spin_inversion, symmetries=model.symmetries)
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    basis = p.product_basis()
    assert not basis.is_built
    basis.unchecked_set_representatives(reps.copy())
    assert basis.is_built and basis.number_states == reps.shape[0]
    assert np.array_equal(basis.states, reps)
    rng = np.random.default_rng(11)
    some = reps[rng.integers(0, reps.shape[0], size=2000)]
    assert np.array_equal(basis.index(some), np.searchsorted(reps, some))
    assert np.all(basis.index(some ^ np.uint64(1)) == np.where(np.isin(some ^ np.uint64(1), reps),
                                                                np.searchsorted(reps, some ^ np.uint64(1)), -1))
    op = ls.Operator(basis, p.expr)
    real = all((2 * g.phase).denominator == 1 for g in model.symmetries.elements)   # (the oracle's matvec is real-only)
    if real:
        x = rng.standard_normal(reps.shape[0])
        y = op.apply_to_state_vector(x)
        want, nnz = oracle.matvec(ob, off, diag, index, x)
        assert np.linalg.norm(y - want) <= 1e-12 * np.linalg.norm(want)
        assert op.count_matrix_elements() == nnz
    # the same basis built on the GPU gives the same product (the lazily computed norms are the build's)
    built = p.product_basis()
    built.build()
    assert np.array_equal(built.states, reps)
    if real:
        again = ls.Operator(built, p.expr).apply_to_state_vector(x)
        assert np.linalg.norm(again - y) <= 1e-14 * np.linalg.norm(y)


# ---- the reference's headline example, as it is written there ----------------------------------------------------------
def test_getting_started_example(oracle):
    """python/example/getting_started.py: NO Hamming weight on the basis (every magnetisation sector the symmetries allow:
    34 states, enumerated by plain index under the projection), ``eigsh`` on the operator, E0 = -18.06178542."""
    import importlib.util
    from pathlib import Path
    import lattice_symmetries_b200 as ls
    path = Path(__file__).resolve().parent.parent / "examples" / "getting_started.py"
    spec = importlib.util.spec_from_file_location("getting_started_example", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    assert abs(module.main(verbose=False) - (-18.06178542)) < 1e-7
    # the same basis against the oracle, state by state
    p = H.chain10_getting_started()
    p.hamming_weight = None
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    basis = ls.SpinBasis(10, None, -1, p.symmetries)
    basis.build()
    assert reps.shape[0] == 34 and np.array_equal(basis.states, reps)
    x = np.random.default_rng(3).standard_normal(34)
    want, _ = oracle.matvec(ob, off, diag, index, x)
    assert np.allclose(ls.Operator(basis, p.expr).apply_to_state_vector(x), want, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("sites,inversion,translation,parity", [(16, 1, 0, 0), (14, -1, 7, 1), (18, None, 2, None)])
def test_projected_basis_without_a_hamming_weight(oracle, sites, inversion, translation, parity):
    """The combination of the reference's example at sizes where the scan spans many 32-candidate words: all 2^n states
    are candidates (plain index, no combinadics), the group projection keeps the representatives of every magnetisation."""
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    model = L.heisenberg_chain(sites)
    syms = L.chain_symmetries(sites, translation, parity)
    p = H.Problem(f"chain{sites}_nohw", sites, model.expression, hamming_weight=None, spin_inversion=inversion,
                  symmetries=syms)
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    basis = p.product_basis()
    basis.build()
    assert np.array_equal(basis.states, reps)
    betas, chars, norms = basis.state_info(reps)
    want_norms = ob.group.state_info(reps)[2]
    assert np.array_equal(betas, reps) and np.array_equal(norms.view(np.uint64), want_norms.view(np.uint64))
    assert np.array_equal(basis.index(reps), np.arange(reps.shape[0]))
    if all((2 * g.phase).denominator == 1 for g in syms.elements):
        x = np.random.default_rng(sites).standard_normal(reps.shape[0])
        want, nnz = oracle.matvec(ob, off, diag, index, x)
        op = ls.Operator(basis, p.expr)
        y = op.apply_to_state_vector(x)
        assert np.linalg.norm(y - want) <= 1e-12 * np.linalg.norm(want) and op.count_matrix_elements() == nnz


# ---- load_yaml_config -> library ----------------------------------------------------------------------------------------
def test_load_yaml_config(oracle, tmp_path):
    from lattice_symmetries_b200.config import load_yaml_config, parse_yaml_file
    path = tmp_path / "chain10.yaml"
    path.write_text(H.CHAIN10_YAML, encoding="utf-8")
    config = load_yaml_config(str(path))
    assert config.hamiltonian is not None and len(config.observables) == 1
    config.basis.build()
    ob, reps, index, off, diag = H.problem_of(parse_yaml_file(path)).oracle_setup(oracle)
    assert np.array_equal(config.basis.states, reps) and reps.shape[0] == 13
    x = np.random.default_rng(5).standard_normal(13)
    want, _ = oracle.matvec(ob, off, diag, index, x)
    assert np.allclose(config.hamiltonian.apply_to_state_vector(x), want, rtol=1e-12, atol=1e-13)
    dense = np.stack([config.hamiltonian.apply_to_state_vector(e) for e in np.eye(13)], axis=1)
    assert np.isclose(np.linalg.eigvalsh(dense)[0], -18.06178542, atol=1e-8)   # python/example/getting_started.py:51


def test_basis_json_round_trip_and_pretty_states():
    """python/lattice_symmetries/__init__.py:248-254, 320-329."""
    import json
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    b = L.heisenberg_chain(10).basis()
    again = ls.Basis.from_json(b.to_json())
    assert isinstance(again, ls.SpinBasis) and json.loads(again.to_json()) == json.loads(b.to_json())
    b.build()
    again.build()
    assert np.array_equal(b.states, again.states)
    assert b.state_to_string(int(b.states[0])) == "|" + format(int(b.states[0]), "010b") + "⟩"
    f = ls.SpinfulFermionBasis(3, (2, 1))
    assert json.loads(f.to_json()) == {"particle": "spinful-fermion", "number_sites": 3, "number_particles": [2, 1]}
    assert ls.Basis.from_json(f.to_json()).to_json() == f.to_json()
    assert f.state_to_string(0b100110) == "|100⟩|110⟩"
    assert json.loads(ls.SpinlessFermionBasis(5, 2).to_json())["number_particles"] == 2


# ---- the reference's golden-file tests, on golden files in its layout -------------------------------------------------------
@pytest.mark.parametrize("stem", ["heisenberg_chain_16_symm", "heisenberg_kagome_18_symm"])
def test_reference_style_golden_files(stem):
    """chapel/test/TestStatesEnumeration.chpl:23-27 (representatives ``==``) and TestMatrixVectorProduct.chpl:15-20, 34-52
    (y within max(1e-13, 1e-11 max(|a|, |b|))) as ``reference_tests.check_states`` / ``check_matvec``: model from the YAML
    file, data from the HDF5 file (tests/golden, written by tests/golden/make_golden.py from the oracle)."""
    from pathlib import Path
    from lattice_symmetries_b200 import hdf5
    from lattice_symmetries_b200 import reference_tests as R
    golden = Path(__file__).parent / "golden"
    model, data = golden / f"{stem}.yaml", golden / f"{stem}.h5"
    lines = []
    out = R.check_states(model, data, log=lines.append)
    assert out.ok and out.details["dim"] == hdf5.File(data).shape("/representatives")[0], lines
    out = R.check_matvec(model, data, log=lines.append)
    assert out.ok and out.details["max_abs_err"] < 1e-12, lines
    out = R.benchmark(model, log=lines.append, repeats=1)
    assert out.details["matrix_elements_per_s"] > 0 and out.details["dim"] == hdf5.File(data).shape("/representatives")[0]


# ---- the program ------------------------------------------------------------------------------------------------------
def test_diagonalize_program_chain10(oracle, tmp_path):
    from lattice_symmetries_b200 import hdf5
    from lattice_symmetries_b200.config import parse_yaml_file
    from lattice_symmetries_b200.diagonalize import diagonalize
    path = tmp_path / "chain10.yaml"
    path.write_text(H.CHAIN10_YAML, encoding="utf-8")
    out = tmp_path / "chain10.h5"
    res = diagonalize(path, out, num_evals=3, eps=1e-10)
    stand_in = H.OracleOperator(oracle, H.problem_of(parse_yaml_file(path)))
    dense = stand_in.dense()
    exact = np.linalg.eigvalsh(dense)
    assert res.dim == 13 and res.converged and not res.reused_representatives
    assert abs(res.eigenvalues[0] - (-18.06178542)) < 1e-8
    assert np.allclose(res.eigenvalues, exact[:3], atol=1e-9)
    assert np.array_equal(hdf5.read_dataset(out, "basis/representatives"), stand_in.reps)
    vecs = hdf5.read_dataset(out, "hamiltonian/eigenvectors")
    assert vecs.shape == (3, 13)
    for e, v in zip(res.eigenvalues, vecs):
        assert np.linalg.norm(dense @ v - e * v) < 1e-8
    assert np.array_equal(hdf5.read_dataset(out, "hamiltonian/eigenvalues"), res.eigenvalues)
    # second run: the representatives come from the file (ls_hs_unchecked_set_representatives on the GPU side)
    again = diagonalize(path, out, num_evals=2, eps=1e-10)
    assert again.reused_representatives and np.allclose(again.eigenvalues, exact[:2], atol=1e-9)


@pytest.mark.parametrize("sites,k", [(16, 2), (20, 3)])
def test_diagonalize_program_with_restarts(oracle, tmp_path, sites, k):
    """dim > Krylov basis: thick restarts on the GPU, eigenpairs against the dense matrix of the oracle's operator."""
    from lattice_symmetries_b200 import hdf5
    from lattice_symmetries_b200 import lattices as L
    from lattice_symmetries_b200.config import parse_yaml_file
    from lattice_symmetries_b200.diagonalize import diagonalize
    path = tmp_path / f"chain{sites}.yaml"
    path.write_text(H.yaml_of_model(L.heisenberg_chain(sites)), encoding="utf-8")
    out = tmp_path / f"chain{sites}.h5"
    res = diagonalize(path, out, num_evals=k, eps=1e-9, max_basis_size=20)
    stand_in = H.OracleOperator(oracle, H.problem_of(parse_yaml_file(path)))
    assert res.dim == stand_in.layout.dim and res.dim > 20 and res.converged
    dense = stand_in.dense()
    exact = np.linalg.eigvalsh(dense)
    assert np.allclose(res.eigenvalues, exact[:k], atol=1e-7 * abs(exact[0]))
    vecs = hdf5.read_dataset(out, "hamiltonian/eigenvectors")
    for e, v in zip(res.eigenvalues, vecs):
        assert np.linalg.norm(dense @ v - e * v) < 1e-6 * abs(e)
    assert np.array_equal(hdf5.read_dataset(out, "basis/representatives"), stand_in.reps)


# ---- block eigensolver over the library's block product ------------------------------------------------------------------
@pytest.mark.parametrize("block", [2, 4])
def test_block_lanczos_on_the_block_product(oracle, block):
    """k lowest pairs with ``block`` vectors per pass over the matrix elements (ls_b200_matvec_block_device) against
    scipy's eigsh on the ORACLE's operator and against the single-vector method."""
    import scipy.sparse.linalg as sla
    import torch
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    from lattice_symmetries_b200.lanczos import lanczos_block_thick_restart, lanczos_thick_restart
    model = L.kagome_heisenberg(18)
    p = H.Problem(model.name, model.number_sites, model.expression, hamming_weight=model.hamming_weight,
                  spin_inversion=model.spin_inversion, symmetries=model.symmetries)
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    basis = p.product_basis()
    basis.build()
    op = ls.Operator(basis, p.expr)
    dim = reps.shape[0]
    mv = lambda v: oracle.matvec(ob, off, diag, index, np.ascontiguousarray(v, dtype=np.float64).reshape(-1))[0]
    want = np.sort(sla.eigsh(sla.LinearOperator((dim, dim), matvec=mv, dtype=np.float64), k=4, which="SA", tol=1e-12)[0])
    res = lanczos_block_thick_restart(op, k=4, block_size=block, basis_size=40, tol=1e-11)
    assert res.converged
    assert np.allclose(res.energies, want, rtol=0, atol=1e-9 * abs(want[0]))
    for i in range(4):
        v = res.eigenvectors[i].contiguous()
        w = torch.zeros_like(v)
        op.matvec_device(v.data_ptr(), w.data_ptr(), sync=True)
        assert float(torch.linalg.vector_norm(w - res.energies[i] * v)) < 1e-8 * abs(want[0])
    single = lanczos_thick_restart(op, k=4, basis_size=40, tol=1e-11)
    assert np.allclose(single.energies, res.energies, rtol=0, atol=1e-9 * abs(want[0]))
