"""GPU parity tests: the CUDA path (through the Python host mirror -> C ABI of
include/lattice_symmetries_b200.h) against the CPU oracle on the same inputs.

Bar: bit-exact for representatives, norms, indices, flags, (beta, coeff, offsets)
rows; 1e-12 relative for matvec output (BASELINE.json north_star); 1e-10 for
ground-state energies.  Run with ``pytest -m gpu`` on a B200.
"""
from __future__ import annotations

import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

MATVEC_RTOL = 1e-12


def _ls():
    import lattice_symmetries_b200 as ls
    return ls


def _rel_err(a, b):
    scale = max(float(np.linalg.norm(b)), 1e-300)
    return float(np.linalg.norm(a - b)) / scale


# ---- problems ------------------------------------------------------------------------
def _model_problem(model) -> H.Problem:
    particle = 0 if model.particle == "spin-1/2" else 1
    return H.Problem(model.name, model.number_sites, model.expression, particle=particle,
                     hamming_weight=model.hamming_weight, number_particles=model.number_particles,
                     spin_inversion=model.spin_inversion, symmetries=model.symmetries)


def _problems():
    from lattice_symmetries_b200 import lattices as L
    return {
        "chain10": H.chain10_getting_started,
        "chain16_symm": lambda: _model_problem(L.heisenberg_chain(16)),
        "chain20_k3": lambda: _model_problem(L.heisenberg_chain(20, translation_sector=3, parity_sector=None,
                                                                spin_inversion=None)),
        "chain24_symm": lambda: _model_problem(L.heisenberg_chain(24)),     # BASELINE configs[0]
        "kagome12_complex": H.kagome12_complex_sector,
        "kagome18_c2": lambda: _model_problem(L.kagome_heisenberg(18)),
        "kagome24_c2v_inv": lambda: _model_problem(L.kagome_heisenberg(24, spin_inversion=1)),
        "ladder_2x8_dm": lambda: _model_problem(L.ladder_dm(8)),            # configs[2] shape, scaled down
        "hubbard_2x4": H.hphi_04_hubbard_square,                            # configs[3] shape, scaled down
        "hubbard_3x3_43": lambda: _model_problem(L.hubbard_square(3, 3, number_particles=(4, 3))),
        "hphi01": H.hphi_01_kagome,
        "hphi02": H.hphi_02_ladder,
        "hphi03": H.hphi_03_hcor,
        "hphi05": H.hphi_05_hubbard_tri,
        # wide states at low filling: the hi-word planes (NP > 32), NP > 48 (no room for (term, sign) in the staged
        # states: row_combine path) and NP = 64, at sizes the oracle enumerates instantly
        "chain40_hw3": lambda: H.Problem("chain40_hw3", 40, L.heisenberg_chain(40).expression, hamming_weight=3,
                                         symmetries=L.heisenberg_chain(40).symmetries),
        "chain56_hw3": lambda: H.Problem("chain56_hw3", 56, L.heisenberg_chain(56).expression, hamming_weight=3,
                                         symmetries=L.heisenberg_chain(56).symmetries),
        "chain64_hw2": lambda: H.Problem("chain64_hw2", 64, L.heisenberg_chain(64).expression, hamming_weight=2,
                                         symmetries=L.heisenberg_chain(64).symmetries),
        "chain12_inv_only": lambda: H.Problem(
            "chain12_inv_only", 12, L.heisenberg_chain(12).expression, hamming_weight=6, spin_inversion=-1),
        "chain10_inv_nohw": lambda: H.Problem(
            "chain10_inv_nohw", 10, L.heisenberg_chain(10).expression, spin_inversion=1),
    }


SYMMETRIC = ["chain10", "chain16_symm", "chain20_k3", "chain24_symm", "kagome12_complex", "kagome18_c2",
             "kagome24_c2v_inv", "ladder_2x8_dm", "chain40_hw3", "chain56_hw3", "chain64_hw2"]
ALL = list(_problems().keys())
REAL_MATVEC = [k for k in ALL if k not in ("ladder_2x8_dm", "kagome12_complex", "chain20_k3")]


@pytest.fixture(scope="module")
def built():
    """name -> (problem, oracle setup, product basis, product operator), cached per module."""
    cache = {}

    def get(name, oracle):
        if name not in cache:
            ls = _ls()
            p = _problems()[name]()
            setup = p.oracle_setup(oracle)
            basis = p.product_basis()
            basis.build()
            op = ls.Operator(basis, p.expr)
            cache[name] = (p, setup, basis, op)
        return cache[name]
    return get


# ---- (a-9) basis construction ----------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_build_matches_oracle(oracle, built, name):
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    got = basis.states
    assert got.dtype == np.uint64
    assert got.shape == reps.shape, (got.shape, reps.shape)
    assert np.array_equal(got, reps)
    assert np.all(got[1:] > got[:-1]) if got.size > 1 else True


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome12_complex", "kagome24_c2v_inv"])
@pytest.mark.parametrize("mode", ["scalar", "bitsliced", "onepass"])
def test_build_kernel_variants_agree(oracle, name, mode, monkeypatch):
    """Both pass-A kernels (bit-sliced plane renaming / scalar Benes walk) emit the oracle's list."""
    monkeypatch.setenv("LS_B200_BUILD", mode)
    p = _problems()[name]()
    reps = p.oracle_basis(oracle).enumerate()
    basis = p.product_basis()
    basis.build()
    assert np.array_equal(basis.states, reps)


@pytest.mark.parametrize("name", SYMMETRIC)
def test_built_norms_bit_exact(oracle, built, name):
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    d_reps, d_norms, count = basis.device_view()
    assert count == reps.shape[0]
    assert np.array_equal(_lib.device_to_numpy(d_reps, count, np.uint64), reps)
    norms = _lib.device_to_numpy(d_norms, count, np.float64)
    _, _, want = ob.group.state_info(reps)
    assert np.array_equal(norms.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("name", ["chain24_symm", "kagome18_c2", "hubbard_2x4", "hphi01"])
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_build_concatenates(oracle, built, name, shards):
    """Contiguous candidate-index shards (one per rank) concatenate to the full sorted list."""
    from lattice_symmetries_b200 import _lib
    from lattice_symmetries_b200.distributed import shard_bounds
    p, (ob, reps, *_), basis, op = built(name, oracle)
    fresh = p.product_basis()
    total = fresh.number_candidates
    parts = []
    for r in range(shards):
        lo, hi = shard_bounds(total, shards, r)
        d_reps, d_norms, count = fresh.build_shard(lo, hi)
        parts.append(_lib.device_to_numpy(d_reps, count, np.uint64))
        _lib.lib.ls_b200_device_free(d_reps)
        if d_norms:
            _lib.lib.ls_b200_device_free(d_norms)
    assert np.array_equal(np.concatenate(parts), reps)


@pytest.mark.parametrize("name", ["chain24_symm", "kagome18_c2", "hubbard_2x4", "hphi01"])
@pytest.mark.parametrize("world", [2, 3])
def test_block_cyclic_build_interleaves(oracle, built, name, world):
    """A rank's block-cyclic share in one call (ls_b200_build_blocks); the ranks' pieces interleave
    block by block to the full sorted list, with norms."""
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, *_), basis, op = built(name, oracle)
    fresh = p.product_basis()
    total = fresh.number_candidates
    size = max(64, (total // (3 * world) + 31) // 32 * 32)   # equal blocks (multiples of 32), ~3 per rank
    plan = [(a, min(total, a + size)) for a in range(0, total, size)]
    pieces = {}
    for rank in range(world):
        mine = plan[rank::world]
        if not mine:
            continue
        d_reps, d_norms, counts = fresh.build_blocks(mine[0][0], size, size * world, len(mine))
        states = _lib.device_to_numpy(d_reps, sum(counts), np.uint64)
        norms = _lib.device_to_numpy(d_norms, sum(counts), np.float64) if d_norms else None
        start = 0
        for k, c in enumerate(counts):
            pieces[rank + k * world] = (states[start:start + c], None if norms is None else norms[start:start + c])
            start += c
        _lib.lib.ls_b200_device_free(d_reps)
        if d_norms:
            _lib.lib.ls_b200_device_free(d_norms)
    assert sorted(pieces) == list(range(len(plan)))
    assert np.array_equal(np.concatenate([pieces[b][0] for b in range(len(plan))]), reps)
    if ob.group is not None and pieces[0][1] is not None:
        want = ob.group.state_info(reps)[2]
        got = np.concatenate([pieces[b][1] for b in range(len(plan))])
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_basis_lists_reference_known_answers():
    """python/test/test_api.py:38-42, python/run_tests.py:102-113."""
    ls = _ls()
    b = ls.SpinBasis(4)
    b.build()
    assert np.array_equal(b.states, np.arange(16, dtype=np.uint64))
    assert np.array_equal(b.index(b.states), np.arange(16))
    b = ls.SpinfulFermionBasis(2)
    b.build()
    assert np.array_equal(b.states, np.arange(16, dtype=np.uint64))
    b = ls.SpinfulFermionBasis(2, 1)
    b.build()
    assert b.states.tolist() == [1, 2, 4, 8]
    b = ls.SpinlessFermionBasis(5, 2)
    b.build()
    assert b.states.tolist() == sorted(x for x in range(32) if bin(x).count("1") == 2)


# ---- (a-1, a-2) per-state kernels ---------------------------------------------------------
@pytest.mark.parametrize("name", SYMMETRIC)
def test_state_info_and_is_representative(oracle, built, name):
    p, (ob, reps, *_), basis, op = built(name, oracle)
    rng = np.random.default_rng(42)
    n = p.number_sites
    hw = p.hamming_weight if p.hamming_weight is not None else n // 2
    states = np.concatenate([H.random_fixed_hamming_states(rng, n, hw, 20000), reps[:20000]])
    betas, chars, norms = basis.state_info(states)
    wb, wc, wn = ob.group.state_info(states)
    assert np.array_equal(betas, wb)
    # States outside the sector have a mathematically-zero stabiliser sum; with complex
    # characters the reference's fp sum leaves +-1e-16 noise there (sqrt -> ~1e-9 or NaN,
    # SURVEY 8a-2).  The CUDA path clamps that band to exactly 0; everything else is bit-exact.
    live = wn > 1e-6
    assert np.array_equal(norms[live].view(np.uint64), wn[live].view(np.uint64))
    assert np.all(norms[~live] == 0.0)
    assert live.sum() >= reps[:20000].shape[0]
    assert np.array_equal(chars[live].view(np.float64).view(np.uint64), wc[live].view(np.float64).view(np.uint64))
    flags, sums = basis.is_representative(states)
    wf, ws = ob.group.is_representative(states)
    assert np.array_equal(flags, wf)
    keep = wf == 1
    assert np.array_equal(sums[keep].view(np.uint64), ws[keep].view(np.uint64))


def test_state_info_empty_and_strided(oracle, built):
    p, (ob, reps, *_), basis, op = built("chain16_symm", oracle)
    betas, chars, norms = basis.state_info(np.zeros(0, dtype=np.uint64))
    assert betas.shape == (0,) and chars.shape == (0,) and norms.shape == (0,)
    b, c, nrm = basis.state_info(int(reps[3]))
    assert b == int(reps[3]) and c == 1.0 + 0j and nrm > 0


def _sanitise_index(idx, dim):
    """kernels/indexing.c:196-215 reads representatives[dim] (one past the end) for needles
    above the largest representative of the last bucket and reports index ``dim`` when the
    stray word happens to equal the needle (reproduced with the reference's own compiled
    source: reps followed in memory by 593 -> index(593) == 13 == dim).  That is undefined
    behaviour, not a contract; the only valid answer for such a needle is -1."""
    idx = idx.copy()
    idx[idx >= dim] = -1
    return idx


# ---- (a-4) state_index ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_state_index(oracle, built, name):
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(7)
    present = reps[rng.integers(0, reps.shape[0], size=min(50000, 4 * reps.shape[0]))]
    absent = present ^ np.uint64(1)  # mostly not in the basis
    junk = rng.integers(0, 2 ** min(63, ob.number_bits), size=1000, dtype=np.uint64)
    needles = np.concatenate([reps[:1], reps[-1:], present, absent, junk])
    got = basis.index(needles)
    want = _sanitise_index(index(needles), reps.shape[0])
    assert np.array_equal(got, want)
    assert np.array_equal(basis.index(reps), np.arange(reps.shape[0]))
    if oracle.ref_available():
        ref = _sanitise_index(oracle.ref_state_index(reps, ob.number_bits, 22, needles), reps.shape[0])
        assert np.array_equal(got, ref)


def test_state_index_crowded_buckets(oracle):
    """Fixed-Hamming-weight representatives pile up behind leading zeros: a basis large
    enough for the prefix table to grow a second level (index.cu sub-tables)."""
    from lattice_symmetries_b200 import lattices as L
    p = _model_problem(L.heisenberg_chain(26))
    basis = p.product_basis()
    basis.build()
    reps = np.asarray(basis.states)
    info = basis.index_info()
    assert info["two_level"] and info["states"] == reps.shape[0] and info["steps"] <= 9, info
    assert np.all(reps[1:] > reps[:-1])
    assert np.array_equal(basis.index(reps), np.arange(reps.shape[0]))
    rng = np.random.default_rng(17)
    absent = reps[rng.integers(0, reps.shape[0], size=20000)] ^ np.uint64(2)
    got = basis.index(absent)
    pos = np.searchsorted(reps, absent)
    pos[pos == reps.shape[0]] = 0
    want = np.where(reps[pos] == absent, pos, -1)
    assert np.array_equal(got, want)
    # an unprojected fixed-weight basis is the extreme case: every candidate is present
    full = H.Problem("chain22_full", 22, L.heisenberg_chain(22).expression, hamming_weight=11).product_basis()
    full.build()
    states = np.asarray(full.states)
    assert np.array_equal(full.index(states), np.arange(states.shape[0]))
    assert np.all(full.index(states[:5000] ^ np.uint64(1)) == -1)


# ---- (a-5, a-6) rows of H ---------------------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_operator_apply(oracle, built, name):
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(3)
    states = reps[rng.integers(0, reps.shape[0], size=min(5000, reps.shape[0]))]
    xs = rng.standard_normal(states.shape[0])
    for scale in (None, xs):
        betas, coeffs, offsets = op.apply_off_diag(states, scale)
        wb, wc, wo = oracle.apply_off_diag(off, states, scale)
        assert np.array_equal(offsets, wo)
        assert np.array_equal(betas, wb)
        assert np.array_equal(coeffs.view(np.float64).view(np.uint64), wc.view(np.float64).view(np.uint64))
        ys = op.apply_diag(states, scale)
        wy = oracle.apply_diag(diag, states, scale)
        assert np.allclose(ys, wy, rtol=1e-15, atol=1e-15)


@pytest.mark.parametrize("particles,expected", [(None, H.HUBBARD2_MATRIX_16), (2, H.HUBBARD2_MATRIX_6)])
def test_hubbard_dense_matrices(particles, expected):
    """python/run_tests.py:128-177 through the vtable's single-state row queries."""
    ls = _ls()
    p = H.hubbard2(particles)
    basis = p.product_basis()
    basis.build()
    op = ls.Operator(basis, p.expr)
    M = H.dense_from_rows(op.apply_diag_to_basis_state, op.apply_off_diag_to_basis_state, basis.states)
    assert np.array_equal(M, expected)
    # and the matvec reproduces the same matrix column by column
    dim = basis.number_states
    cols = np.stack([op.apply_to_state_vector(e) for e in np.eye(dim)], axis=1)
    assert np.allclose(cols, expected, atol=1e-14)


# ---- (a-7, a-8) matvec ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", REAL_MATVEC)
def test_matvec_matches_oracle(oracle, built, name):
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(42)
    x = rng.standard_normal(reps.shape[0])
    x /= np.linalg.norm(x)
    y = op.apply_to_state_vector(x)
    want, nnz = oracle.matvec(ob, off, diag, index, x)
    assert _rel_err(y, want) < MATVEC_RTOL, _rel_err(y, want)
    assert np.allclose(y, want, rtol=1e-11, atol=1e-13)  # chapel/test/TestMatrixVectorProduct.chpl:15-16
    assert op.count_matrix_elements() == nnz


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv"])
def test_matvec_scalar_variant_agrees(oracle, built, name, monkeypatch):
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(reps.shape[0])
    y0 = op.apply_to_state_vector(x)
    monkeypatch.setenv("LS_B200_MATVEC", "scalar")
    y1 = op.apply_to_state_vector(x)
    assert _rel_err(y1, y0) < MATVEC_RTOL


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv", "kagome18_c2", "chain40_hw3",
                                  "chain56_hw3", "chain64_hw2"])
@pytest.mark.parametrize("chunk", [None, "4096"])
def test_matvec_pipeline_variants_agree(oracle, built, name, chunk, monkeypatch):
    """Fused (canonicalise + rank + gather in one kernel, default) vs the three-kernel
    pipeline, also with tiny row chunks so that warps straddle chunk ends."""
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(6)
    x = rng.standard_normal(reps.shape[0])
    y0 = op.apply_to_state_vector(x)
    if chunk is not None:
        monkeypatch.setenv("LS_B200_MV_CHUNK", chunk)
        y1 = op.apply_to_state_vector(x)
        assert np.array_equal(y1, y0)  # chunking never changes the summation order
    for variant in ("unfused", "split", "fused", "scalar"):
        monkeypatch.setenv("LS_B200_MATVEC", variant)
        y2 = op.apply_to_state_vector(x)
        assert _rel_err(y2, y0) < MATVEC_RTOL, variant


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv", "kagome18_c2", "chain40_hw3"])
@pytest.mark.parametrize("chunk,bits", [(None, None), ("8192", "6"), ("300000", "64")])
def test_matvec_sorted_ranking_is_bit_identical(oracle, built, name, chunk, bits, monkeypatch):
    """Sorted ranking (the large-footprint path: representatives radix-sorted by their leading bits before the index
    search, values scattered back to CSR order) changes the ORDER of the lookups only: y is bit-identical, with
    chunks cut by element count (tiny, uneven, single) and any number of sort bits."""
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    x = np.random.default_rng(16).standard_normal(reps.shape[0])
    monkeypatch.setenv("LS_B200_MV_SORT", "0")
    y0 = op.apply_to_state_vector(x)
    monkeypatch.setenv("LS_B200_MV_SORT", "1")
    if chunk is not None:
        monkeypatch.setenv("LS_B200_MV_CHUNK", chunk)
        monkeypatch.setenv("LS_B200_MV_SORT_BITS", bits)
    y1 = op.apply_to_state_vector(x)
    assert np.array_equal(y1, y0)
    want = oracle.matvec(ob, off, diag, index, x)[0]
    assert _rel_err(y1, want) < MATVEC_RTOL
    # device-resident entry point on a row range
    from lattice_symmetries_b200 import _lib
    dim = reps.shape[0]
    d_x = _lib.DeviceArray.from_numpy(x)
    d_y = _lib.DeviceArray(dim, np.float64)
    lo, hi = dim // 3, dim - dim // 5
    op.matvec_device(d_x.ptr, d_y.ptr, lo, hi, sync=True)
    assert np.array_equal(d_y.numpy()[:hi - lo], y0[lo:hi])


@pytest.mark.parametrize("name", ["ladder_2x8_dm", "kagome12_complex"])
def test_matvec_sorted_ranking_complex(oracle, built, name, monkeypatch):
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    d_x = _lib.DeviceArray.from_numpy(np.random.default_rng(8).standard_normal(dim) + 1j * np.random.default_rng(9).standard_normal(dim))
    d_y = _lib.DeviceArray(dim, np.complex128)
    monkeypatch.setenv("LS_B200_MV_SORT", "0")
    op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)
    y0 = d_y.numpy().copy()
    monkeypatch.setenv("LS_B200_MV_SORT", "1")
    monkeypatch.setenv("LS_B200_MV_CHUNK", "4096")
    op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)
    assert np.array_equal(d_y.numpy(), y0)


@pytest.mark.parametrize("name", ["ladder_2x8_dm", "kagome12_complex"])
def test_matvec_complex_pipeline_variants_agree(oracle, built, name, monkeypatch):
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    rng = np.random.default_rng(8)
    d_x = _lib.DeviceArray.from_numpy(rng.standard_normal(dim) + 1j * rng.standard_normal(dim))
    d_y = _lib.DeviceArray(dim, np.complex128)
    op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)
    y0 = d_y.numpy().copy()
    for variant in ("unfused", "split", "fused", "scalar"):
        monkeypatch.setenv("LS_B200_MATVEC", variant)
        op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)
        assert _rel_err(d_y.numpy(), y0) < MATVEC_RTOL, variant


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv", "hubbard_2x4", "chain12_inv_only"])
def test_block_matvec_equals_single_vectors(oracle, built, name, monkeypatch):
    """Extension (the reference halts on numVectors != 1): k vectors through one pass == k products, bit for bit;
    also on a row range with strided device vectors and with tiny chunks."""
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    rng = np.random.default_rng(12)
    X = rng.standard_normal((3, dim))
    want = np.stack([op.apply_to_state_vector(x) for x in X])
    assert np.array_equal(op.apply_to_state_vector(X), want)
    monkeypatch.setenv("LS_B200_MV_CHUNK", "4096")
    assert np.array_equal(op.apply_to_state_vector(X), want)
    monkeypatch.delenv("LS_B200_MV_CHUNK")
    lo, hi = dim // 4, dim - dim // 5
    pad = 7
    d_x = _lib.DeviceArray.from_numpy(np.concatenate([np.concatenate([x, np.zeros(pad)]) for x in X]))
    d_y = _lib.DeviceArray(3 * (hi - lo + pad), np.float64)
    op.matvec_block_device(3, d_x.ptr, dim + pad, d_y.ptr, hi - lo + pad, lo, hi, sync=True)
    got = d_y.numpy().reshape(3, hi - lo + pad)[:, :hi - lo]
    assert np.array_equal(got, want[:, lo:hi])


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv", "chain56_hw3", "hubbard_2x4",
                                  "ladder_2x8_dm"])
def test_phased_matvec_equals_plain(oracle, built, name, monkeypatch):
    """canonicalise (phase 1) + apply (phase 2) == the plain product, also when phase 1 ran for an earlier vector,
    when it is skipped (phase 2 catches up) and with several chunks."""
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    cplx = name == "ladder_2x8_dm"
    dtype = np.complex128 if cplx else np.float64
    rng = np.random.default_rng(21)
    lo, hi = dim // 5, dim - dim // 7
    for chunk in (None, "4096"):
        if chunk is not None:
            monkeypatch.setenv("LS_B200_MV_CHUNK", chunk)
        for trial in range(3):
            x = rng.standard_normal(dim) + (1j * rng.standard_normal(dim) if cplx else 0)
            d_x = _lib.DeviceArray.from_numpy(x.astype(dtype))
            d_y = _lib.DeviceArray(hi - lo, dtype)
            d_z = _lib.DeviceArray(hi - lo, dtype)
            op.matvec_device(d_x.ptr, d_y.ptr, lo, hi, complex_vectors=cplx, sync=True)
            if trial != 1:
                op.matvec_device_phase(1, 0, 0, lo, hi, complex_vectors=cplx)
            op.matvec_device_phase(2, d_x.ptr, d_z.ptr, lo, hi, complex_vectors=cplx)
            _lib.lib.ls_b200_matvec_sync()
            _lib.check_error()
            assert np.array_equal(d_z.numpy(), d_y.numpy()), (chunk, trial)


@pytest.mark.parametrize("name", ["chain16_symm", "chain24_symm", "kagome24_c2v_inv", "hubbard_2x4", "chain56_hw3"])
@pytest.mark.parametrize("chunk", [None, "4096"])
def test_matvec_pinned_host_buffers(oracle, built, name, chunk, monkeypatch):
    """The reference-facing call with PINNED host buffers takes the overlapped route (x uploads under the
    canonicalise phase, y drains chunk by chunk): same bits as with pageable buffers."""
    import ctypes as C
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    lib = _lib.lib
    x = np.random.default_rng(31).standard_normal(dim)
    want = op.apply_to_state_vector(x)
    if chunk is not None:
        monkeypatch.setenv("LS_B200_MV_CHUNK", chunk)
    hx, hy = lib.ls_b200_host_malloc(8 * dim), lib.ls_b200_host_malloc(8 * dim)
    try:
        np.frombuffer((C.c_double * dim).from_address(hx), dtype=np.float64)[:] = x
        mv = lib.ls_hs_internal_get_chpl_kernels().contents.matrix_vector_product
        for _ in range(2):  # the second call reuses the cached element counts / buffers
            np.frombuffer((C.c_double * dim).from_address(hy), dtype=np.float64)[:] = np.nan
            mv(C.byref(op._payload), 1, C.cast(hx, _lib.f64_p), C.cast(hy, _lib.f64_p))
            _lib.check_error()
            got = np.frombuffer((C.c_double * dim).from_address(hy), dtype=np.float64).copy()
            assert np.array_equal(got, want)
    finally:
        lib.ls_b200_host_free(hx)
        lib.ls_b200_host_free(hy)


@pytest.mark.parametrize("name", ["chain16_symm", "kagome18_c2", "hubbard_2x4"])
def test_matvec_device_row_ranges(oracle, built, name):
    """Device-resident entry point on contiguous row shards == host-pointer entry point."""
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    rng = np.random.default_rng(11)
    x = rng.standard_normal(dim)
    want = op.apply_to_state_vector(x)
    d_x = _lib.DeviceArray.from_numpy(x)
    parts = []
    bounds = [0, dim // 3, dim // 3, (2 * dim) // 3 + 1, dim]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        d_y = _lib.DeviceArray(hi - lo, np.float64)
        op.matvec_device(d_x.ptr, d_y.ptr, lo, hi, sync=True)
        parts.append(d_y.numpy())
    assert np.array_equal(np.concatenate(parts), want)


@pytest.mark.parametrize("name", ["ladder_2x8_dm", "kagome12_complex", "chain20_k3", "chain16_symm"])
def test_matvec_complex_hermitian_and_dense(oracle, built, name):
    """Complex128 vectors are an extension with no reference oracle:
    validate against a dense matrix assembled from the oracle's per-state
    primitives with the rule H[j,i] = chi v sign n_j / n_i (BatchedOperator.chpl:207-253)."""
    from lattice_symmetries_b200 import _lib
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    dim = reps.shape[0]
    rng = np.random.default_rng(9)
    x = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    # dense reference (columns)
    betas, coeffs, offsets = oracle.apply_off_diag(off, reps)
    rb, rc, rn = ob.group.state_info(betas)
    _, _, n_alpha = ob.group.state_info(reps)
    j = index(rb)
    col = np.repeat(np.arange(dim), np.diff(offsets))
    live = rn > 0
    assert np.all(j[live] >= 0)
    val = np.where(live, coeffs * rc * rn / n_alpha[col], 0)
    want = np.zeros(dim, dtype=np.complex128)
    np.add.at(want, j[live], (val * x[col])[live])
    want += oracle.apply_diag(diag, reps) * x
    d_x = _lib.DeviceArray.from_numpy(x)
    d_y = _lib.DeviceArray(dim, np.complex128)
    op.matvec_device(d_x.ptr, d_y.ptr, complex_vectors=True, sync=True)
    y = d_y.numpy()
    assert _rel_err(y, want) < MATVEC_RTOL, _rel_err(y, want)
    # Hermiticity: <u, H v> == conj(<v, H u>)
    u = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    d_u = _lib.DeviceArray.from_numpy(u)
    op.matvec_device(d_u.ptr, d_y.ptr, complex_vectors=True, sync=True)
    hu = d_y.numpy()
    assert abs(np.vdot(u, y) - np.conj(np.vdot(x, hu))) < 1e-10 * (np.linalg.norm(u) * np.linalg.norm(y))


def test_matvec_invalid_sector_raises(oracle):
    """DistributedMatrixVector.chpl:127-135: an operator that leaves the symmetry
    sector halts the reference; here it surfaces through ls_hs_error."""
    ls = _ls()
    from lattice_symmetries_b200 import lattices as L
    m = L.heisenberg_chain(12)
    basis = m.basis()
    basis.build()
    # sigma^x changes the Hamming weight: every image leaves the Sz = 0 basis
    bad = ls.Operator(basis, ls.Expr("σˣ₀", sites=[[i] for i in range(12)]))
    x = np.ones(basis.number_states)
    with pytest.raises(RuntimeError, match="invalid index"):
        bad.apply_to_state_vector(x)
    for variant in ("unfused", "split", "fused", "scalar"):
        with pytest.MonkeyPatch.context() as mp:
            mp.setenv("LS_B200_MATVEC", variant)
            with pytest.raises(RuntimeError, match="invalid index"):
                bad.apply_to_state_vector(x)
    # the library stays usable afterwards
    good = m.operator(basis)
    assert np.isfinite(good.apply_to_state_vector(x)).all()


def test_matvec_input_validation(oracle, built):
    p, setup, basis, op = built("chain16_symm", oracle)
    with pytest.raises(TypeError):
        op.apply_to_state_vector(np.zeros(basis.number_states, dtype=np.float32))
    with pytest.raises(ValueError):
        op.apply_to_state_vector(np.zeros(basis.number_states + 1))


# ---- energies (pin build + index + matvec jointly) ----------------------------------------------
def test_chain10_ground_state_energy(oracle, built):
    """python/example/getting_started.py:49-51."""
    import scipy.sparse.linalg as sla
    p, setup, basis, op = built("chain10", oracle)
    assert basis.number_states == 13
    dim = basis.number_states
    Hm = np.stack([op.apply_to_state_vector(e) for e in np.eye(dim)], axis=1)
    assert np.allclose(Hm, Hm.T, atol=1e-13)
    assert np.isclose(np.linalg.eigvalsh(Hm)[0], -18.06178542, atol=1e-8)
    w = sla.eigsh(op, k=1, which="SA", tol=1e-12)[0]
    assert np.isclose(w[0], -18.06178542, atol=1e-8)


@pytest.mark.parametrize("name", ["hphi01", "hphi02", "hphi03", "hubbard_2x4", "hphi05"])
def test_hphi_energies(oracle, built, name):
    """test/0N_*/HPhi/output/zvo_energy.dat:1 via eigsh on the GPU operator."""
    import scipy.sparse.linalg as sla
    p, setup, basis, op = built(name, oracle)
    w = sla.eigsh(op, k=1, which="SA", tol=1e-12)[0]
    assert np.isclose(w[0], p.energy, rtol=0, atol=1e-9), (w[0], p.energy)


def test_lanczos_matches_eigsh(oracle, built):
    from lattice_symmetries_b200.lanczos import lanczos_ground_state
    import scipy.sparse.linalg as sla
    p, setup, basis, op = built("chain24_symm", oracle)
    e0 = lanczos_ground_state(op, max_iters=200, tol=1e-12, seed=1).energy
    w = sla.eigsh(op, k=1, which="SA", tol=1e-12)[0]
    assert abs(e0 - w[0]) < 1e-10 * abs(w[0])


# ---- golden fixtures (tests/golden, generated by tests/golden/make_golden.py) -------------------
@pytest.mark.parametrize("name", ["chain16_symm", "kagome18_c2", "kagome24_c2v_inv", "ladder_2x8_dm", "kagome12_complex",
                                  "chain12_inv_only", "hubbard_2x4"])
def test_projected_rows_of_H(oracle, built, name):
    """ls_b200_operator_apply_off_diag_projected: representatives and offsets bit-exact, coefficients
    chi c n(beta) / n(alpha) (BatchedOperator.chpl:207-253) to rounding, indices consistent with the basis; and the
    rows, summed up, reproduce the matrix-vector product."""
    p, (ob, reps, index, off, diag), basis, op = built(name, oracle)
    rng = np.random.default_rng(31)
    sel = np.sort(rng.choice(reps.shape[0], size=min(reps.shape[0], 500), replace=False))
    alphas = reps[sel]
    got_reps, got_c, got_off, got_idx = op.apply_off_diag_projected(alphas, with_indices=True)
    betas, coeffs, offsets = oracle.apply_off_diag(off, alphas)
    assert np.array_equal(got_off, offsets)
    col = np.repeat(np.arange(alphas.shape[0]), np.diff(offsets))
    if ob.c.has_permutation_symmetries:
        rb, rc, rn = ob.group.state_info(betas)
        na = ob.group.state_info(alphas)[2]
        want_c = np.where(rn > 0, rc * (coeffs * rn / na[col]), 0)
    elif p.spin_inversion:
        mask = np.uint64((1 << p.number_sites) - 1)
        flipped = betas ^ mask
        rb = np.minimum(betas, flipped)
        want_c = np.where(flipped < betas, float(p.spin_inversion), 1.0) * coeffs
    else:
        rb, want_c = betas, coeffs
    assert np.array_equal(got_reps, rb)
    assert np.allclose(got_c, want_c, rtol=1e-14, atol=1e-15)
    assert np.array_equal(got_idx, index(rb))
    # columns of the projected operator assemble to H: y = sum_i x_i (column i), restricted to the sampled columns
    if not np.iscomplexobj(want_c) or np.all(np.abs(want_c.imag) < 1e-14):
        x = np.zeros(reps.shape[0])
        x[sel] = rng.standard_normal(sel.shape[0])
        y = oracle.apply_diag(diag, reps) * x if diag.n else np.zeros(reps.shape[0])
        live = got_idx >= 0
        np.add.at(y, got_idx[live], (got_c.real * x[sel][col])[live])
        assert _rel_err(y, op.apply_to_state_vector(x)) < MATVEC_RTOL


def test_lanczos_thick_restart_matches_eigsh(oracle, built):
    """k lowest eigenpairs (Diagonalize.chpl:166-177 numEvals) against scipy's eigsh on the ORACLE's matvec --
    an independent operator implementation on a problem the plain Lanczos test does not use."""
    import scipy.sparse.linalg as sla
    from lattice_symmetries_b200.lanczos import lanczos_ground_state, lanczos_thick_restart
    p, (ob, reps, index, off, diag), basis, op = built("kagome18_c2", oracle)
    dim = reps.shape[0]
    mv = lambda v: oracle.matvec(ob, off, diag, index, np.ascontiguousarray(v, dtype=np.float64).reshape(-1))[0]
    want = np.sort(sla.eigsh(sla.LinearOperator((dim, dim), matvec=mv, dtype=np.float64), k=4, which="SA", tol=1e-12)[0])
    res = lanczos_thick_restart(op, k=4, basis_size=32, tol=1e-11)
    assert res.converged
    assert np.allclose(res.energies, want, rtol=0, atol=1e-9 * abs(want[0]))
    # residuals of the returned vectors, computed with the product itself
    import torch
    for i in range(4):
        v = res.eigenvectors[i]
        w = torch.zeros_like(v)
        op.matvec_device(v.data_ptr(), w.data_ptr(), sync=True)
        assert float(torch.linalg.vector_norm(w - res.energies[i] * v)) < 1e-8 * abs(want[0])
    single = lanczos_ground_state(op, tol=1e-12, energy_tol=1e-13)
    assert abs(single.energy - want[0]) < 1e-9 * abs(want[0])


def test_lanczos_thick_restart_complex_sector(oracle, built):
    """complex128 vectors (momentum sector with complex characters): lowest eigenvalues against a dense
    diagonalisation of the matrix assembled from the oracle's per-state primitives."""
    import torch
    from lattice_symmetries_b200.lanczos import lanczos_thick_restart
    p, (ob, reps, index, off, diag), basis, op = built("kagome12_complex", oracle)
    dim = reps.shape[0]
    betas, coeffs, offsets = oracle.apply_off_diag(off, reps)
    rb, rc, rn = ob.group.state_info(betas)
    na = ob.group.state_info(reps)[2]
    col = np.repeat(np.arange(dim), np.diff(offsets))
    j = index(rb)
    live = rn > 0
    Hm = np.zeros((dim, dim), dtype=np.complex128)
    np.add.at(Hm, (j[live], col[live]), (coeffs * rc * rn / na[col])[live])
    Hm += np.diag(oracle.apply_diag(diag, reps))
    assert np.allclose(Hm, Hm.conj().T, atol=1e-12)
    want = np.linalg.eigvalsh(Hm)[:3]
    res = lanczos_thick_restart(op, k=3, basis_size=30, tol=1e-11, dtype=torch.complex128)
    assert res.converged and np.allclose(res.energies, want, rtol=0, atol=1e-9 * abs(want[0]))


def test_c_caller_built_against_the_reference_header(oracle, tmp_path):
    """tests/abi_harness.c knows only the reference's kernels/lattice_symmetries_types.h; built on the CPU tier
    (tests/test_abi.py) and linked with the drop-in library, it fills the structs like the Haskell host does
    (FFI.hs:92-143), builds the basis with ls_hs_build_representatives and multiplies through
    ls_hs_internal_get_chpl_kernels()->matrix_vector_product.  Representatives bit-exact, y to 1e-12 vs the oracle."""
    import struct
    import subprocess
    from pathlib import Path
    from lattice_symmetries_b200.expr import compile_terms
    from lattice_symmetries_b200 import lattices as L
    harness = Path(__file__).resolve().parent / "_build" / "abi_harness"
    if not harness.exists():
        if not Path("/root/reference/kernels").is_dir():
            pytest.skip("tests/_build/abi_harness has not been built (it needs the reference header: run the CPU tests first)")
        import test_abi
        test_abi.build_harness()
    n, hw, inv = 14, 7, -1
    model = L.heisenberg_chain(n, symmetric=False)
    p = H.Problem("chain14_inv", n, model.expression, hamming_weight=hw, spin_inversion=inv)
    ob, reps, index, off, diag = p.oracle_setup(oracle)
    ts = compile_terms(model.expression, n)
    tables = [[t for t in ts if t.x != 0], [t for t in ts if t.x == 0]]
    blob = struct.pack("<5i", n, hw, inv, len(tables[0]), len(tables[1]))
    for tab in tables:
        blob += b"".join(struct.pack("<2d", t.v.real, t.v.imag) for t in tab)
        for k in "mlrxs":
            blob += b"".join(struct.pack("<Q", getattr(t, k)) for t in tab)
    (tmp_path / "problem.bin").write_bytes(blob)
    out = subprocess.run([str(harness), "run", str(tmp_path / "problem.bin"), str(tmp_path / "out.bin")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = (tmp_path / "out.bin").read_bytes()
    dim = struct.unpack("<Q", raw[:8])[0]
    got_reps = np.frombuffer(raw, dtype=np.uint64, count=dim, offset=8)
    x = np.frombuffer(raw, dtype=np.float64, count=dim, offset=8 + 8 * dim)
    y = np.frombuffer(raw, dtype=np.float64, count=dim, offset=8 + 16 * dim)
    assert np.array_equal(got_reps, reps)
    want = oracle.matvec(ob, off, diag, index, x)[0]
    assert _rel_err(y, want) < MATVEC_RTOL


def test_golden_fixtures():
    import json
    from pathlib import Path
    ls = _ls()
    golden = Path(__file__).parent / "golden"
    from golden import make_golden as G
    for name in G.CASES:
        data = np.load(golden / f"{name}.npz")
        p = G.CASES[name]()
        basis = p.product_basis()
        basis.build()
        assert np.array_equal(basis.states, data["representatives"]), name
        op = ls.Operator(basis, p.expr)
        if "y" in data.files:
            y = op.apply_to_state_vector(data["x"])
            assert _rel_err(y, data["y"]) < MATVEC_RTOL, name
