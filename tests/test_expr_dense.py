"""The expression compiler against an independent dense evaluation.

``expr.py`` (parser + product / merge of non-branching terms: Expr.hs:313-345, NonbranchingTerm.hs:24-58,
Generator.hs:229-247) produces the term tables that BOTH the oracle and the CUDA path consume, so a mistake there is
invisible to every parity test; only the reference's known answers (Hubbard matrices, energies) pin it.  Here random
expressions are evaluated a second way -- every primitive as an explicit 2^n x 2^n matrix, products as matrix products
-- and compared with the matrix the compiled terms generate (T|a> = v (-1)^{|a & s|} [a & m == r] |a ^ x>).
Conventions of the reference: bit i set = spin up / site occupied; sigma^+ = |up><down|; S<op> = sigma<op> / 2 for EVERY
op, sigma^+- included (Expr.hs:332-333); spinful fermions: up sites in bits [0, n), down in [n, 2n), Jordan-Wigner string
over the lower bits (Basis.hs:621-628, Generator.hs:244-247)."""
from __future__ import annotations

import numpy as np
import pytest

from lattice_symmetries_b200.expr import Expr, compile_terms

SUB = "₀₁₂₃₄₅₆₇₈₉"


def _sub(i: int) -> str:
    return "".join(SUB[int(d)] for d in str(i))


def dense_from_terms(terms, bits: int) -> np.ndarray:
    dim = 1 << bits
    M = np.zeros((dim, dim), dtype=np.complex128)
    for t in terms:
        for a in range(dim):
            if (a & t.m) == t.r:
                M[a ^ t.x, a] += t.v * (-1) ** bin(a & t.s).count("1")
    return M


def _site_matrix(bits: int, bit: int, kind: str, string: bool) -> np.ndarray:
    dim = 1 << bits
    M = np.zeros((dim, dim), dtype=np.complex128)
    for a in range(dim):
        occupied = (a >> bit) & 1
        sign = (-1) ** bin(a & ((1 << bit) - 1)).count("1") if string else 1
        if kind == "raise" and not occupied:
            M[a | (1 << bit), a] = sign
        elif kind == "lower" and occupied:
            M[a & ~(1 << bit), a] = sign
        elif kind == "z":
            M[a, a] = 1 if occupied else -1
        elif kind == "n":
            M[a, a] = occupied
    return M


def spin_primitive(n: int, prefix: str, op: str, site: int):
    """(text, matrix)"""
    scale = 0.5 if prefix == "S" else 1.0
    plus, minus, z = (_site_matrix(n, site, k, False) for k in ("raise", "lower", "z"))
    matrix = {"x": plus + minus, "y": -1j * (plus - minus), "z": z, "+": plus, "-": minus}[op] * scale
    glyph = {"x": "ˣ", "y": "ʸ", "z": "ᶻ", "+": "⁺", "-": "⁻"}[op]
    return f"{prefix}{glyph}{_sub(site)}", matrix


def fermion_primitive(n: int, spinful: bool, op: str, spin, site: int):
    bit = site + (n if spin == 1 else 0)
    bits = 2 * n if spinful else n
    kind = {"c†": "raise", "c": "lower", "n": "n"}[op]
    arrow = {None: "", 0: "↑", 1: "↓"}[spin]
    return f"{op}{_sub(site)}{arrow}", _site_matrix(bits, bit, kind, kind != "n")


def random_expression(rng, primitive, max_terms=4, max_factors=3):
    text, total = [], None
    for k in range(int(rng.integers(1, max_terms + 1))):
        coeff = round(float(rng.uniform(-2, 2)), 3)
        factors = [primitive(rng) for _ in range(int(rng.integers(1, max_factors + 1)))]
        matrix = coeff * np.eye(factors[0][1].shape[0], dtype=np.complex128)
        for _, f in factors:          # "A B" = the operator product: B acts first
            matrix = matrix @ f
        sign = "-" if coeff < 0 else ("+" if k else "")
        text.append(f"{sign} {abs(coeff)} " + " ".join(name for name, _ in factors))
        total = matrix if total is None else total + matrix
    return " ".join(text).strip(), total


@pytest.mark.parametrize("seed", range(8))
def test_spin_expressions(seed):
    rng = np.random.default_rng(seed)
    n = 3
    prim = lambda r: spin_primitive(n, "σS"[int(r.integers(2))], "xyz+-"[int(r.integers(5))], int(r.integers(n)))
    for _ in range(25):
        text, want = random_expression(rng, prim)
        got = dense_from_terms(compile_terms(Expr(text), n), n)
        assert np.allclose(got, want, atol=1e-12), text


@pytest.mark.parametrize("seed", range(6))
def test_spinless_fermion_expressions(seed):
    rng = np.random.default_rng(100 + seed)
    n = 3
    prim = lambda r: fermion_primitive(n, False, ["c†", "c", "n"][int(r.integers(3))], None, int(r.integers(n)))
    for _ in range(25):
        text, want = random_expression(rng, prim)
        got = dense_from_terms(compile_terms(Expr(text), n), n)
        assert np.allclose(got, want, atol=1e-12), text


@pytest.mark.parametrize("seed", range(6))
def test_spinful_fermion_expressions(seed):
    rng = np.random.default_rng(200 + seed)
    n = 2
    prim = lambda r: fermion_primitive(n, True, ["c†", "c", "n"][int(r.integers(3))], int(r.integers(2)), int(r.integers(n)))
    for _ in range(25):
        text, want = random_expression(rng, prim)
        got = dense_from_terms(compile_terms(Expr(text), n), 2 * n)
        assert np.allclose(got, want, atol=1e-12), text


def test_sites_expansion_and_algebra():
    """``sites`` rows replace the distinct indices in ascending order (Expr.hs:270-289); + - * scale adjoint."""
    n = 4
    bond = "σˣ₀ σˣ₁ + σʸ₀ σʸ₁ + σᶻ₀ σᶻ₁"
    e = Expr(bond, sites=[[0, 1], [2, 3], [3, 0]])
    explicit = Expr("σˣ₀ σˣ₁ + σʸ₀ σʸ₁ + σᶻ₀ σᶻ₁ + σˣ₂ σˣ₃ + σʸ₂ σʸ₃ + σᶻ₂ σᶻ₃ + σˣ₃ σˣ₀ + σʸ₃ σʸ₀ + σᶻ₃ σᶻ₀")
    A = dense_from_terms(compile_terms(e, n), n)
    assert np.allclose(A, dense_from_terms(compile_terms(explicit, n), n))
    assert np.allclose(A, A.conj().T)
    a, b = Expr("0.7 σ⁺₀ σᶻ₂"), Expr("(1.5 + 0.5im) σʸ₁ σ⁻₃")
    Ma, Mb = (dense_from_terms(compile_terms(x, n), n) for x in (a, b))
    assert np.allclose(dense_from_terms(compile_terms(a + b, n), n), Ma + Mb)
    assert np.allclose(dense_from_terms(compile_terms(a - b, n), n), Ma - Mb)
    assert np.allclose(dense_from_terms(compile_terms(a * b, n), n), Ma @ Mb)
    assert np.allclose(dense_from_terms(compile_terms(b.scale(2 - 1j), n), n), (2 - 1j) * Mb)
    assert np.allclose(dense_from_terms(compile_terms(b.adjoint(), n), n), Mb.conj().T)
    with pytest.raises(ValueError):
        Expr(bond, sites=[[0, 1, 2]])
