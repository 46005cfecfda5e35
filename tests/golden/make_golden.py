"""Generates the golden fixtures in this directory (run in the build container,
where /root/reference is mounted):

    python tests/golden/make_golden.py

For every case: the sorted representatives of the basis, a seeded input vector
x and y = H x.  The numbers come from the CPU oracle (oracle/ls_oracle.c), whose
state_index / operator-application pieces are cross-checked here against the
reference's own C sources compiled from /root/reference (oracle/_ref/libref.so
= kernels/indexing.c + kernels/reference.c) before anything is written; the
orbit arithmetic (Halide generator, unbuildable here) is pinned by the known
answers in tests/test_oracle.py.  The GPU box has no /root/reference: the
``-m gpu`` tests compare the CUDA path with these files.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import helpers as H  # noqa: E402


def _model(model) -> H.Problem:
    particle = 0 if model.particle == "spin-1/2" else 1
    return H.Problem(model.name, model.number_sites, model.expression, particle=particle,
                     hamming_weight=model.hamming_weight, number_particles=model.number_particles,
                     spin_inversion=model.spin_inversion, symmetries=model.symmetries)


def _lattices():
    from lattice_symmetries_b200 import lattices
    return lattices


CASES = {
    "chain10_getting_started": H.chain10_getting_started,
    "chain16_symm": lambda: _model(_lattices().heisenberg_chain(16)),
    "chain20_symm": lambda: _model(_lattices().heisenberg_chain(20)),
    "kagome12_complex_sector": H.kagome12_complex_sector,           # representatives only
    "kagome18_c2": lambda: _model(_lattices().kagome_heisenberg(18)),
    "hphi04_hubbard_square": H.hphi_04_hubbard_square,
    "hphi01_kagome9": H.hphi_01_kagome,
}


# The same numbers in the REFERENCE's golden-file layout (chapel/test/TestStatesEnumeration.chpl:23-25: /representatives
# u64[dim]; chapel/test/TestMatrixVectorProduct.chpl:7-11: /x, /y f64[1, dim]) next to a model file in its YAML schema.
H5_CASES = {
    "heisenberg_chain_16_symm": ("chain16_symm", lambda: _lattices().heisenberg_chain(16)),
    "heisenberg_kagome_18_symm": ("kagome18_c2", lambda: _lattices().kagome_heisenberg(18)),
}


def write_h5_cases(out: Path) -> None:
    from lattice_symmetries_b200 import hdf5
    from lattice_symmetries_b200.config import parse_yaml_file
    from lattice_symmetries_b200.expr import compile_terms
    for stem, (npz_name, make) in H5_CASES.items():
        model = make()
        data = np.load(out / f"{npz_name}.npz")
        (out / f"{stem}.yaml").write_text(H.yaml_of_model(model), encoding="utf-8")
        back = parse_yaml_file(out / f"{stem}.yaml")
        table = lambda e: sorted((t.m, t.r, t.x, t.s, t.l, complex(t.v)) for t in compile_terms(e, model.number_sites))
        assert table(back.hamiltonian) == table(model.expression), stem
        assert np.array_equal(back.model.symmetries.permutations(), model.symmetries.permutations()), stem
        hdf5.create(out / f"{stem}.h5", {"/representatives": hdf5.DatasetSpec(data["representatives"].shape, np.uint64),
                                         "/x": hdf5.DatasetSpec((1, data["x"].shape[0]), np.float64),
                                         "/y": hdf5.DatasetSpec((1, data["y"].shape[0]), np.float64)}, align=8)
        hdf5.write_rows(out / f"{stem}.h5", "/representatives", data["representatives"])
        hdf5.write_rows(out / f"{stem}.h5", "/x", data["x"][None, :])
        hdf5.write_rows(out / f"{stem}.h5", "/y", data["y"][None, :])
        print(f"{stem}.h5 / .yaml: dim={data['representatives'].shape[0]}")


def main() -> None:
    from oracle import ls_oracle as oracle
    oracle.build()
    out = Path(__file__).resolve().parent
    for name, make in CASES.items():
        p = make()
        b, reps, index, off, diag = p.oracle_setup(oracle)
        data = {"representatives": reps}
        # sin(-2 pi / 2) = -1.2e-16 in the reference too (Group.hs:115-116): real up to rounding
        real = p.symmetries is None or bool(np.all(np.abs(p.symmetries.characters()[1]) < 1e-9))
        if oracle.ref_available():
            rng = np.random.default_rng(0)
            needles = np.concatenate([reps, reps ^ np.uint64(1), rng.integers(0, 2 ** b.number_bits, 512, dtype=np.uint64)])
            assert np.array_equal(index(needles), oracle.ref_state_index(reps, b.number_bits, 22, needles)), name
            wb, wc, wo = oracle.ref_apply_off_diag(off, diag, b.number_bits, reps[:2048])
            gb, gc, go = oracle.apply_off_diag(off, reps[:2048])
            assert np.array_equal(wb, gb) and np.array_equal(wc, gc) and np.array_equal(wo, go), name
        if real:
            rng = np.random.default_rng(42)
            x = rng.standard_normal(reps.shape[0])
            x /= np.linalg.norm(x)
            y, nnz = oracle.matvec(b, off, diag, index, x)
            data.update(x=x, y=y, nnz=np.int64(nnz))
        np.savez_compressed(out / f"{name}.npz", **data)
        print(f"{name}: dim={reps.shape[0]} real={real}")
    write_h5_cases(out)


if __name__ == "__main__":
    main()
