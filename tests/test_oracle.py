"""Pins the CPU oracle (oracle/ls_oracle.c) before anything trusts it:

* against the reference's own C sources compiled in place (oracle/_ref/libref.so:
  kernels/indexing.c, kernels/reference.c) -- bit-exact;
* against every known answer the reference's tests hold for this path
  (SURVEY 8c): Benes vectors, basis lists, Hubbard dense matrices, the
  chain-10 energy, the five HPhi energies;
* against a pure-Python restatement of kernels/generator.cpp on small groups.

No GPU needed.
"""
import numpy as np
import pytest

from lattice_symmetries_b200.expr import Expr, compile_terms
from lattice_symmetries_b200.symmetry import (Symmetries, Symmetry, apply_benes, benes_network,
                                             permute_bits_naive)
from lattice_symmetries_b200 import lattices

import helpers as H


# ---- Benes networks: haskell/test/LatticeSymmetries/BenesSpec.hs:34-43 -------------------
def test_benes_known_vectors():
    m, s = benes_network([0, 1, 2])
    assert apply_benes(m, s, 0b100) == 0b100
    m, s = benes_network([1, 2, 0])
    assert apply_benes(m, s, 0b100) == 0b010
    assert apply_benes(m, s, 0b101) == 0b110
    v3 = [1, 2, 3, 0, 5, 6, 7, 4, 9, 10, 11, 8]
    m, s = benes_network(v3)
    assert apply_benes(m, s, 0b100111010010) == 0b110011100001
    assert permute_bits_naive(v3, 0b100111010010) == 0b110011100001


def test_benes_random_permutations():
    rng = np.random.default_rng(42)
    for n in [1, 2, 3, 5, 8, 13, 24, 31, 32, 33, 36, 42, 63, 64]:
        for _ in range(5):
            p = rng.permutation(n).tolist()
            m, s = benes_network(p)
            assert len(s) == len(m)
            for _ in range(20):
                x = int(rng.integers(0, 2**63, dtype=np.uint64)) & ((1 << n) - 1)
                assert apply_benes(m, s, x) == permute_bits_naive(p, x)


# ---- state_info / is_representative vs a pure-Python reading of generator.cpp --------------
def _python_state_info(perms, chars, inv, nbits, x):
    """kernels/generator.cpp:24-54, 77-140 with naive permutations."""
    flip = (1 << nbits) - 1
    r, c, n = x, 1.0 + 0j, 0.0
    for p, ch in zip(perms, chars):
        y = permute_bits_naive(p, x)
        if y < r:
            r, c = y, ch
        if y == x:
            n += ch.real
        if inv:
            yf = y ^ flip
            if yf < r:
                r, c = yf, inv * ch
            if yf == x:
                n += inv * ch.real
    return r, c, np.sqrt(n / ((2 if inv else 1) * len(perms)))


@pytest.mark.parametrize("case", ["chain10", "chain12_k1", "kagome12", "kagome12_c6v_inv"])
def test_oracle_state_info_matches_python(oracle, case):
    if case == "chain10":
        p = H.chain10_getting_started()
        syms, inv, n, hw = p.symmetries, -1, 10, 5
    elif case == "chain12_k1":
        syms, inv, n, hw = lattices.chain_symmetries(12, 1, None), 0, 12, 6
    elif case == "kagome12":
        p = H.kagome12_complex_sector()
        syms, inv, n, hw = p.symmetries, 0, 12, 6
    else:
        m = lattices.kagome_heisenberg(12, spin_inversion=1)
        syms, inv, n, hw = m.symmetries, 1, 12, 6
    g = oracle.Group.from_symmetries(syms, n, inv)
    rng = np.random.default_rng(1)
    xs = H.random_fixed_hamming_states(rng, n, hw, 200)
    betas, chars, norms = g.state_info(xs)
    flags, sums = g.is_representative(xs)
    perms = [list(s._perm) for s in syms.elements]
    re, im = syms.characters()
    ch = [complex(a, b) for a, b in zip(re, im)]
    for i, x in enumerate(xs.tolist()):
        r, c, nrm = _python_state_info(perms, ch, inv, n, x)
        assert betas[i] == r
        assert chars[i] == c
        assert norms[i] == nrm or (np.isnan(norms[i]) and np.isnan(nrm))
        # generator.cpp:201-253
        is_rep = all(permute_bits_naive(p, x) >= x and (not inv or (permute_bits_naive(p, x) ^ ((1 << n) - 1)) >= x)
                     for p in perms)
        if abs(sums[i]) > 1e-9:  # away from the FP-noise band of complex sectors
            assert bool(flags[i]) == (is_rep and sums[i] > 0)
        if flags[i]:
            assert betas[i] == x


# ---- basis lists: python/test/test_api.py:38-42, python/run_tests.py:102-113 ---------------
def test_basis_lists(oracle):
    b = oracle.Basis(4, 0, 4, None, None, oracle.Group.from_symmetries(Symmetries([]), 4, None))
    states = b.enumerate()
    assert states.tolist() == list(range(16))
    assert oracle.Index(states, 4)(states).tolist() == list(range(16))
    assert oracle.Basis(2, 1).enumerate().tolist() == list(range(16))
    assert oracle.Basis(2, 1, 1, None).enumerate().tolist() == [1, 2, 4, 8]
    # (n_up, n_down) product enumeration, StatesEnumeration.chpl:290-326
    assert oracle.Basis(2, 1, 2, 1).enumerate().tolist() == [0b0101, 0b0110, 0b1001, 0b1010]


def test_combinadics_roundtrip(oracle):
    L = oracle.lib()
    import math
    for n, k in [(4, 2), (10, 5), (24, 12), (36, 18), (42, 21), (64, 32), (64, 1), (7, 0)]:
        total = math.comb(n, k)
        for idx in {i for i in (0, 1, total // 3, total - 1) if i < total}:
            s = L.oracle_fixed_hamming_index_to_state(idx, k)
            assert bin(s).count("1") == k and s < (1 << n)
            assert L.oracle_fixed_hamming_state_to_index(s) == idx
    # consecutive indices are Gosper successors
    s = L.oracle_fixed_hamming_index_to_state(1000, 5)
    t = s | (s - 1)
    nxt = (t + 1) | (((~t & (t + 1)) - 1) >> ((s & -s).bit_length()))
    assert L.oracle_fixed_hamming_index_to_state(1001, 5) == nxt


# ---- Hubbard dense matrices: python/run_tests.py:116-177 --------------------------------------
@pytest.mark.parametrize("particles,expected", [(None, H.HUBBARD2_MATRIX_16), (2, H.HUBBARD2_MATRIX_6)])
def test_hubbard_dense_matrices(oracle, particles, expected):
    p = H.hubbard2(particles)
    states = p.oracle_basis(oracle).enumerate()
    off, diag = p.terms(oracle)

    def diag_row(ket):
        return float(oracle.apply_diag(diag, np.array([ket], dtype=np.uint64))[0])

    def off_row(ket):
        b, c, _ = oracle.apply_off_diag(off, np.array([ket], dtype=np.uint64))
        return list(zip(c.tolist(), b.tolist()))

    assert np.array_equal(H.dense_from_rows(diag_row, off_row, states), expected)
    if oracle.ref_available():  # the reference's own reference.c gives the same rows

        def diag_row_ref(ket):
            return float(oracle.ref_apply_diag(off, diag, 4, np.array([ket], dtype=np.uint64))[0])

        def off_row_ref(ket):
            b, c, _ = oracle.ref_apply_off_diag(off, diag, 4, np.array([ket], dtype=np.uint64))
            return list(zip(c.tolist(), b.tolist()))

        assert np.array_equal(H.dense_from_rows(diag_row_ref, off_row_ref, states), expected)


def test_operator_apply_counts(oracle):
    """python/test/test_api.py:64-69"""
    ts = compile_terms(Expr("1.0 σᶻ₀ σᶻ₁ + 2.0 σ⁺₀ σ⁻₁ + 2.0 σ⁻₀ σ⁺₁"), 2)
    off = oracle.Terms([t for t in ts if t.x])
    b, c, offsets = oracle.apply_off_diag(off, np.array([1], dtype=np.uint64))
    assert len(b) == 1 and b[0] == 2 and c[0] == 2.0


# ---- oracle vs the compiled reference C, bit-exact ---------------------------------------------
def test_oracle_index_matches_reference_c(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (reference tree absent)")
    rng = np.random.default_rng(7)
    for nbits, count, prefix in [(24, 5000, 22), (36, 20000, 22), (36, 20000, 10), (40, 3, 26), (16, 1, 22),
                                 (63, 1000, 22)]:
        reps = np.unique(rng.integers(0, 2**nbits, size=count, dtype=np.uint64))
        needles = np.concatenate([reps[rng.integers(0, len(reps), size=2000)],
                                  rng.integers(0, 2**nbits, size=2000, dtype=np.uint64)])
        a = oracle.Index(reps, nbits, prefix)(needles)
        b = oracle.ref_state_index(reps, nbits, prefix, needles)
        assert np.array_equal(a, b)
        present = np.isin(needles, reps)
        assert np.array_equal(a >= 0, present)
        assert np.array_equal(reps[a[present]], needles[present])


@pytest.mark.parametrize("maker", [H.hphi_01_kagome, H.hphi_02_ladder, H.hphi_04_hubbard_square,
                                   lambda: H.Problem("dm", 8, lattices.ladder_dm(4).expression, hamming_weight=4)])
def test_oracle_apply_matches_reference_c(oracle, maker):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (reference tree absent)")
    p = maker()
    b = p.oracle_basis(oracle)
    states = b.enumerate()[:3000]
    off, diag = p.terms(oracle)
    xs = np.random.default_rng(3).normal(size=len(states))
    for x in (None, xs):
        assert np.array_equal(oracle.apply_diag(diag, states, x),
                              oracle.ref_apply_diag(off, diag, b.number_bits, states, x))
        o = oracle.apply_off_diag(off, states, x)
        r = oracle.ref_apply_off_diag(off, diag, b.number_bits, states, x)
        for u, v in zip(o, r):
            assert np.array_equal(u, v)


# ---- energies: getting_started.py:51 and test/0N_*/HPhi/output/zvo_energy.dat:1 ----------------
def test_chain10_energy(oracle):
    p = H.chain10_getting_started()
    e, dim = H.oracle_ground_state_energy(oracle, p)
    assert dim == 13
    assert len(p.symmetries.elements) == 20
    assert np.isclose(e, p.energy)  # the reference's own assertion (np.isclose defaults)
    assert abs(e - (-18.06178541796816)) < 1e-10  # SURVEY headline fact 7


@pytest.mark.parametrize("maker,dim", [(H.hphi_01_kagome, 126), (H.hphi_02_ladder, 64), (H.hphi_03_hcor, 256),
                                       (H.hphi_04_hubbard_square, 4900), (H.hphi_05_hubbard_tri, 4096)])
def test_hphi_energies(oracle, maker, dim):
    p = maker()
    e, d = H.oracle_ground_state_energy(oracle, p)
    assert d == dim
    assert e == pytest.approx(p.energy, rel=1e-6)  # python/test/test_api.py:141-153


def test_symmetric_spectrum_is_subset_of_full(oracle):
    """Projection check: every eigenvalue of the symmetry-reduced chain-12
    matrix (all sectors of T, P, inversion) appears in the unsymmetrised one,
    and the dimensions add up."""
    n = 12
    full = H.Problem("c12", n, lattices.heisenberg_chain(n, symmetric=False).expression, hamming_weight=6)
    b, reps, index, off, diag = full.oracle_setup(oracle)
    Hfull = np.stack([oracle.matvec(b, off, diag, index, e)[0] for e in np.eye(len(reps))], axis=1)
    wfull = np.linalg.eigvalsh(Hfull)
    total = 0
    for k in range(n):
        syms = lattices.chain_symmetries(n, k, None)
        p = H.Problem("c12k", n, full.expr, hamming_weight=6, symmetries=syms)
        b, reps, index, off, diag = p.oracle_setup(oracle)
        total += len(reps)
        if k in (0, n // 2):  # real sectors: the real-valued matvec is the full story
            Hk = np.stack([oracle.matvec(b, off, diag, index, e)[0] for e in np.eye(len(reps))], axis=1)
            for w in np.linalg.eigvalsh(Hk):
                assert np.min(np.abs(wfull - w)) < 1e-9
    assert total == len(wfull)


def test_c1_dimension(oracle):
    """chain-24 symm has 28,968 representatives, |G| = 48."""
    m = lattices.heisenberg_chain(24)
    assert len(m.symmetries.elements) == 48
    p = H.Problem("c1", 24, m.expression, hamming_weight=12, spin_inversion=1, symmetries=m.symmetries)
    b = p.oracle_basis(oracle)
    reps = b.enumerate_range(b.min_state(), b.max_state())
    assert len(reps) == 28968
    assert np.all(reps[1:] > reps[:-1])
    # the threaded chunked enumeration equals the sequential one on a prefix
    seq = b.enumerate()
    assert np.array_equal(seq, reps)
