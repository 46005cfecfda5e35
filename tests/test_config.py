"""CPU tests of the YAML model reader (``config.py``) and of the exact-diagonalisation program (``diagonalize.py``).

The reader is pinned on the reference's OWN model files, read where they lie under /root/reference (skipped when the
tree is absent, as on the GPU box): all 32 of them parse; ``chapel/data/heisenberg_chain_24_symm.yaml`` -- BASELINE.json
configs[0] -- yields bit for bit the term tables and the group the benchmark synthesises; the five ``test/0N_*``
models reproduce the HPhi energies stored next to them (``HPhi/output/zvo_energy.dat``) through the oracle.  The
program (YAML in, HDF5 out: chapel/src/Diagonalize.chpl:258-333) runs here on a stand-in operator backed by the oracle;
on the GPU the same program runs on the library (tests/test_yaml_config.py)."""
from __future__ import annotations

import glob
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from helpers import CHAIN10_YAML, OracleOperator, problem_of
from lattice_symmetries_b200 import hdf5
from lattice_symmetries_b200.config import parse_config, parse_yaml_file
from lattice_symmetries_b200.diagonalize import diagonalize
from lattice_symmetries_b200.expr import compile_terms

REFERENCE = Path("/root/reference")
needs_reference = pytest.mark.skipif(not REFERENCE.exists(), reason="the reference tree is not present")

def term_table(expr, n):
    return sorted((t.m, t.r, t.x, t.s, t.l, complex(t.v)) for t in compile_terms(expr, n))


# ---- the reader --------------------------------------------------------------------------------------------------------
def test_inline_document(tmp_path):
    import yaml
    p = parse_config(yaml.safe_load(CHAIN10_YAML), name="chain10")
    m = p.model
    assert (m.number_sites, m.hamming_weight, m.spin_inversion, m.particle) == (10, 5, -1, "spin-1/2")
    assert len(m.symmetries) == 2 and len(m.symmetries.elements) == 20 and not m.symmetries.is_empty
    assert len(p.observables) == 1 and p.extra == {"number_vectors": 2, "output": "chain10.h5"}
    want = H.chain10_getting_started()
    assert term_table(p.hamiltonian, 10) == term_table(want.expr, 10)
    path = tmp_path / "chain10.yaml"
    path.write_text(CHAIN10_YAML, encoding="utf-8")
    q = parse_yaml_file(path)
    assert q.model.name == "chain10" and term_table(q.hamiltonian, 10) == term_table(want.expr, 10)


@pytest.mark.parametrize("document,message", [
    ({}, "basis"),
    ({"basis": {"number_spins": 4, "spin_inversion": 2}}, "spin_inversion"),
    ({"basis": {"number_spins": 4, "hamming_weight": 5}}, "hamming_weight"),
    ({"basis": {"number_spins": "four"}}, "integer"),
    ({"basis": {"particle": "boson", "number_sites": 4}}, "particle"),
    ({"basis": {"number_spins": 4, "symmetries": [{"permutation": [1, 0, 2], "sector": 0}]}}, "length"),
    ({"basis": {"number_spins": 4, "symmetries": [{"permutation": [1, 0, 3, 2]}]}}, "sector"),
    ({"basis": {"number_spins": 4}, "hamiltonian": {"terms": []}}, "terms"),
    ({"basis": {"number_spins": 4}, "hamiltonian": {"terms": [{"sites": [[0, 1]]}]}}, "expression"),
    ({"basis": {"number_spins": 4}, "hamiltonian": {"terms": [{"expression": "σᶻ₀", "particle": "spinful-fermion"}]}},
     "particle"),
    ({"basis": {"particle": "spinful-fermion", "number_sites": 4, "number_particles": [1, 2, 3]}}, "number_particles"),
    ({"basis": {"number_spins": 4}, "observables": {"terms": []}}, "observables"),
])
def test_schema_errors(document, message):
    with pytest.raises(ValueError, match=message):
        parse_config(document)


def test_fermion_headers():
    p = parse_config({"basis": {"particle": "spinful-fermion", "number_sites": 3, "number_particles": [2, 1]},
                      "hamiltonian": {"terms": [{"expression": "n₀↑ n₀↓", "sites": [[0], [1], [2]]}]}})
    assert p.model.particle == "spinful-fermion" and p.model.number_particles == (2, 1)
    p = parse_config({"basis": {"particle": "spinful-fermion", "number_sites": 3, "number_particles": 3}})
    assert p.model.number_particles == 3 and p.hamiltonian is None
    p = parse_config({"basis": {"particle": "spinless-fermion", "number_sites": 5, "number_particles": 2}})
    assert (p.model.particle, p.model.number_sites, p.model.number_particles) == ("spinless-fermion", 5, 2)


def test_basis_json_and_pretty_states():
    """``basisHeaderToJSON`` / ``basisHeaderFromJSON`` (Basis.hs:289-319) and the state pretty-printer (Basis.hs:103-138)."""
    import json
    from lattice_symmetries_b200 import lattices as L
    from lattice_symmetries_b200.config import basis_header, state_to_string
    m = L.heisenberg_chain(10)
    text = json.dumps(basis_header("spin-1/2", 10, 5, -1, m.symmetries))
    back = parse_config({"basis": json.loads(text)}).model
    assert (back.number_sites, back.hamming_weight, back.spin_inversion) == (10, 5, -1)
    assert np.array_equal(back.symmetries.permutations(), m.symmetries.permutations())
    assert json.loads(text)["symmetries"][0] == {"permutation": [1, 2, 3, 4, 5, 6, 7, 8, 9, 0], "sector": 0}
    plain = json.loads(json.dumps(basis_header("spin-1/2", 4)))
    assert plain == {"particle": "spin-1/2", "number_spins": 4, "hamming_weight": None, "spin_inversion": None,
                     "symmetries": []}
    assert parse_config({"basis": plain}).model.symmetries is None
    for occupation in (None, 3, (2, 1)):
        h = json.loads(json.dumps(basis_header("spinful-fermion", 3, number_particles=occupation)))
        assert parse_config({"basis": h}).model.number_particles == occupation
    h = basis_header("spinless-fermion", 5, number_particles=2)
    assert h == {"particle": "spinless-fermion", "number_sites": 5, "number_particles": 2}
    assert state_to_string(0b0101, 4) == "|0101⟩" and state_to_string(1, 1) == "|1⟩"
    assert state_to_string(0b100110, 6, spinful=True) == "|100⟩|110⟩"


def _reference_models():
    files = sorted(glob.glob(str(REFERENCE / "chapel/data/*.yaml")) + glob.glob(str(REFERENCE / "test/*/hamiltonian.yaml")))
    return [f for f in files if "basis:" in Path(f).read_text(encoding="utf-8")]


@needs_reference
def test_every_model_file_of_the_reference_parses():
    files = _reference_models()
    assert len(files) == 32
    for f in files:
        p = parse_yaml_file(f)
        m = p.model
        assert p.hamiltonian is not None and m.number_sites >= 4
        bits = m.number_sites * (2 if m.particle == "spinful-fermion" else 1)
        terms = compile_terms(p.hamiltonian, m.number_sites)
        assert terms and all(0 <= t.x < (1 << bits) and (t.r & ~t.m) == 0 for t in terms), f
        if m.symmetries is not None:
            assert len(m.symmetries.elements) >= 2


@needs_reference
def test_chain24_symm_is_the_benchmark_configuration():
    """BASELINE.json configs[0] = chapel/data/heisenberg_chain_24_symm.yaml: the synthesised model IS that file."""
    from lattice_symmetries_b200 import lattices as L
    p = parse_yaml_file(REFERENCE / "chapel/data/heisenberg_chain_24_symm.yaml")
    mine = L.heisenberg_chain(24)
    m = p.model
    assert (m.number_sites, m.hamming_weight, m.spin_inversion) == (mine.number_sites, mine.hamming_weight, mine.spin_inversion)
    assert term_table(p.hamiltonian, 24) == term_table(mine.expression, 24)
    assert len(m.symmetries.elements) == len(mine.symmetries.elements) == 48
    assert np.array_equal(m.symmetries.permutations(), mine.symmetries.permutations())
    for a, b in zip(m.symmetries.characters(), mine.symmetries.characters()):
        assert np.array_equal(a, b)


@needs_reference
@pytest.mark.parametrize("folder,restated", [
    ("01_spin_kagome", H.hphi_01_kagome), ("02_spin_ladder_DM", H.hphi_02_ladder), ("03_spin_hcor", H.hphi_03_hcor),
    ("04_hubbard_square", H.hphi_04_hubbard_square), ("05_hubbard_tri", H.hphi_05_hubbard_tri)])
def test_hphi_models_and_energies(oracle, folder, restated):
    """test/0N_*/hamiltonian.yaml read in place; the energy is the one HPhi stored next to it."""
    p = parse_yaml_file(REFERENCE / "test" / folder / "hamiltonian.yaml")
    want = restated()
    n = p.model.number_sites
    assert term_table(p.hamiltonian, n) == term_table(want.expr, n)
    text = (REFERENCE / "test" / folder / "HPhi/output/zvo_energy.dat").read_text()
    energy = float(text.split()[1])
    assert abs(energy - want.energy) < 1e-12
    e0, dim = H.oracle_ground_state_energy(oracle, problem_of(p))
    assert abs(e0 - energy) < 1e-8, (e0, energy, dim)


# ---- the program, on a stand-in operator backed by the oracle ----------------------------------------------------------
def test_run_options_come_from_the_model_file(tmp_path):
    from lattice_symmetries_b200.diagonalize import run_options
    path = tmp_path / "chain10.yaml"
    path.write_text(CHAIN10_YAML, encoding="utf-8")
    assert run_options(path) == {"output": "chain10.h5", "num_evals": 2, "max_basis_size": 0, "max_block_size": 1}
    assert run_options(path, output="x.h5", num_evals=5, max_basis_size=30, max_block_size=4) == \
        {"output": "x.h5", "num_evals": 5, "max_basis_size": 30, "max_block_size": 4}
    if REFERENCE.exists():
        got = run_options(REFERENCE / "chapel/data/heisenberg_square_6x6.yaml")
        assert got == {"output": "data/heisenberg_square_6x6.h5", "num_evals": 2, "max_basis_size": 20, "max_block_size": 4}


def _run(oracle, yaml_path, out, **kw):
    made = {}

    def factory(parsed, cached):
        op = OracleOperator(oracle, problem_of(parsed))
        if cached is not None:
            assert np.array_equal(cached, op.reps)
        made["op"] = op
        return op, op.reps, None

    lines = []
    res = diagonalize(yaml_path, out, operator_factory=factory, log=lines.append, **kw)
    return res, made["op"], lines


def test_diagonalize_chain10(oracle, tmp_path):
    path = tmp_path / "chain10.yaml"
    path.write_text(CHAIN10_YAML, encoding="utf-8")
    out = tmp_path / "out" / "chain10.h5"
    res, op, lines = _run(oracle, path, out, num_evals=3, eps=1e-10)
    assert res.dim == 13 and res.converged and not res.reused_representatives
    dense = op.dense()
    exact = np.linalg.eigvalsh(dense)
    assert abs(res.eigenvalues[0] - (-18.06178542)) < 1e-8
    assert np.allclose(res.eigenvalues, exact[:3], atol=1e-9)
    assert any("Hilbert space dimension: 13" in s for s in lines)
    with hdf5.File(out) as f:
        assert f.datasets() == ["/basis/representatives", "/hamiltonian/eigenvalues", "/hamiltonian/eigenvectors",
                                "/hamiltonian/residuals"]
        assert np.array_equal(f.read("basis/representatives"), op.reps)
        assert np.array_equal(f.read("hamiltonian/eigenvalues"), res.eigenvalues)
        assert np.array_equal(f.read("hamiltonian/residuals"), res.residuals)
        vecs = f.read("hamiltonian/eigenvectors")
    assert vecs.shape == (3, 13) and vecs.dtype == np.float64
    assert np.allclose(vecs @ vecs.T, np.eye(3), atol=1e-9)
    for e, v in zip(res.eigenvalues, vecs):
        assert np.linalg.norm(dense @ v - e * v) < 1e-8
    # a second run finds basis/representatives in the output file and reuses them (makeBasisStates)
    again, _, lines = _run(oracle, path, out, num_evals=2, eps=1e-10)
    assert again.reused_representatives and np.allclose(again.eigenvalues, exact[:2], atol=1e-9)
    assert any("read from the output file" in s for s in lines)
    with hdf5.File(out) as f:
        assert f.shape("hamiltonian/eigenvectors") == (2, 13)
    with pytest.raises(ValueError):
        _run(oracle, path, out, num_evals=0)
    # kMaxBlockSize > 1: the block method over the block product
    blocked, op, _ = _run(oracle, path, tmp_path / "blocked.h5", num_evals=3, eps=1e-10, max_block_size=3)
    assert blocked.converged and np.allclose(blocked.eigenvalues, exact[:3], atol=1e-9) and op.block_products > 0


@needs_reference
@pytest.mark.parametrize("name,k", [("heisenberg_kagome_12", 1), ("heisenberg_kagome_12_symm", 2), ("heisenberg_square_4x4", 2),
                                    ("heisenberg_chain_10", 4)])
def test_diagonalize_reference_models(oracle, tmp_path, name, k):
    """The reference's own inputs (chapel/data): restarts are exercised (dim > Krylov basis), eigenpairs checked
    against the dense matrix of the oracle's operator."""
    out = tmp_path / f"{name}.h5"
    res, op, _ = _run(oracle, REFERENCE / "chapel/data" / f"{name}.yaml", out, num_evals=k, eps=1e-9, max_basis_size=20,
                      max_block_size=2 if name.endswith("4x4") else 1)
    assert res.converged and res.dim == op.layout.dim
    if res.dim <= 1500:
        dense = op.dense()
        exact = np.linalg.eigvalsh(dense)
        assert np.allclose(res.eigenvalues, exact[:k], atol=1e-7 * max(1.0, abs(exact[0])))
        vecs = hdf5.read_dataset(out, "hamiltonian/eigenvectors")
        for e, v in zip(res.eigenvalues, vecs):
            assert np.linalg.norm(dense @ v - e * v) < 1e-6 * max(1.0, abs(e))
    assert np.array_equal(hdf5.read_dataset(out, "basis/representatives"), op.reps)


# ---- golden files in the reference's layout ------------------------------------------------------------------------------
@pytest.mark.parametrize("stem,npz", [("heisenberg_chain_16_symm", "chain16_symm"), ("heisenberg_kagome_18_symm", "kagome18_c2")])
def test_reference_style_golden_files_hold_the_oracle_numbers(oracle, stem, npz):
    """tests/golden/<model>.yaml + <model>.h5 (``/representatives`` u64[dim], ``/x`` ``/y`` f64[1, dim]: the inputs of
    chapel/test/TestStatesEnumeration.chpl and TestMatrixVectorProduct.chpl) against the oracle, freshly computed."""
    golden = Path(__file__).parent / "golden"
    p = problem_of(parse_yaml_file(golden / f"{stem}.yaml"))
    b, reps, index, off, diag = p.oracle_setup(oracle)
    with hdf5.File(golden / f"{stem}.h5") as f:
        assert f.datasets() == ["/representatives", "/x", "/y"]
        assert f.shape("/x") == f.shape("/y") == (1, reps.shape[0])
        assert np.array_equal(f.read("/representatives"), reps)
        x, y = f.read("/x")[0], f.read("/y")[0]
    want, _ = oracle.matvec(b, off, diag, index, x)
    assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)   # (the oracle adds atomically, like the reference: order varies)
    data = np.load(golden / f"{npz}.npz")
    assert np.array_equal(data["representatives"], reps) and np.array_equal(data["x"], x) and np.array_equal(data["y"], y)


# ---- the program on several ranks: one output file, every rank writes its rows ------------------------------------------
def _worker_diagonalize(rank, world, port, folder, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ls_oracle as oracle
        from test_lanczos_cpu import ShardedDenseOperator
        oracle.build()
        path = Path(folder) / "model.yaml"
        target = Path(folder) / "result.h5"
        made = {}

        def factory(parsed, cached):
            assert cached is None                     # several ranks always build
            whole = OracleOperator(oracle, problem_of(parsed))
            dense = whole.dense()
            dim = dense.shape[0]
            bounds = [0, dim // 3, dim] if world == 2 else [0, dim // 3, dim // 3, dim]
            made["dense"], made["reps"], made["bounds"] = dense, whole.reps, bounds
            op = ShardedDenseOperator(dense, rank, world, bounds)
            return op, whole.reps[bounds[rank]:bounds[rank + 1]], None

        res = diagonalize(path, target, num_evals=3, eps=1e-10, max_basis_size=24, operator_factory=factory,
                          barrier=dist.barrier)
        ok = bool(res.converged and res.dim == made["dense"].shape[0])
        exact = np.linalg.eigvalsh(made["dense"])
        ok &= bool(np.allclose(res.eigenvalues, exact[:3], atol=1e-8 * abs(exact[0])))
        # after the final barrier the one file holds every rank's rows
        ok &= bool(np.array_equal(hdf5.read_dataset(target, "basis/representatives"), made["reps"]))
        vecs = hdf5.read_dataset(target, "hamiltonian/eigenvectors")
        for e, v in zip(res.eigenvalues, vecs):
            ok &= bool(np.linalg.norm(made["dense"] @ v - e * v) < 1e-7 * abs(exact[0]))
        ok &= bool(np.array_equal(hdf5.read_dataset(target, "hamiltonian/eigenvalues"), res.eigenvalues))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_diagonalize_across_ranks_under_gloo(tmp_path, world):
    import socket
    import torch.multiprocessing as mp
    from lattice_symmetries_b200 import lattices as L
    (tmp_path / "model.yaml").write_text(H.yaml_of_model(L.heisenberg_chain(16)), encoding="utf-8")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        out = manager.dict()
        procs = [ctx.Process(target=_worker_diagonalize, args=(r, world, port, str(tmp_path), out)) for r in range(world)]
        for pr in procs:
            pr.start()
        for pr in procs:
            pr.join(timeout=300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        assert dict(out) == {r: True for r in range(world)}


# ---- the reference's test programs (chapel/test, chapel/benchmark) on a stand-in backend --------------------------------
def _oracle_backend(oracle):
    import torch

    def backend(parsed):
        op = OracleOperator(oracle, problem_of(parsed))

        def matvec(x):
            y = torch.zeros(op.layout.dim, dtype=torch.float64)
            op.matvec(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)), y)
            return y.numpy()

        count = lambda: oracle.matvec(op.b, op.off, op.diag, op.index, np.zeros(op.layout.dim))[1]
        return op.reps, matvec, count, 0.0, op
    return backend


def test_reference_test_programs(oracle, tmp_path):
    from lattice_symmetries_b200 import reference_tests as R
    golden = Path(__file__).parent / "golden"
    model, data = golden / "heisenberg_chain_16_symm.yaml", golden / "heisenberg_chain_16_symm.h5"
    backend = _oracle_backend(oracle)
    lines = []
    out = R.check_states(model, data, backend=backend, log=lines.append)
    assert out.ok and out.details["dim"] == 257 and len(lines) == 1 and float(lines[0]) >= 0
    lines = []
    out = R.check_matvec(model, data, backend=backend, log=lines.append)
    assert out.ok and lines[0] == "true" and out.details["max_abs_err"] < 1e-13
    lines = []
    out = R.benchmark(model, backend=backend, log=lines.append, repeats=1)
    assert lines[0] == "Hilbert space dimension: 257" and out.details["matrix_elements"] > 257
    # a golden file that disagrees: one representative and one entry of y changed
    reps = hdf5.read_dataset(data, "/representatives").copy()
    y = hdf5.read_dataset(data, "/y").copy()
    reps[5] += 1
    y[0, 7] += 1e-9
    bad = tmp_path / "bad.h5"
    hdf5.write_file(bad, {"/representatives": reps, "/x": hdf5.read_dataset(data, "/x"), "/y": y})
    lines = []
    out = R.check_states(model, bad, backend=backend, log=lines.append)
    assert not out.ok and lines[0].startswith("at index 5: ")
    lines = []
    out = R.check_matvec(model, bad, backend=backend, log=lines.append)
    assert not out.ok and lines[0] == "false" and lines[1].startswith("at 7: ")
    assert np.array_equal(R.approx_equal([1.0, 1.0, 0.0], [1.0 + 5e-12, 1.0 + 2e-11, 5e-14]), [True, False, True])
    short = tmp_path / "short.h5"
    hdf5.write_file(short, {"/representatives": reps[:100]})
    assert not R.check_states(model, short, backend=backend, log=lines.append).ok
