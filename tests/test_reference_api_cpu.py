"""The CPU-checkable half of the reference's own Python tests (python/test/test_api.py, python/run_tests.py) on this
package's host-side mirror of the API: ``Symmetry`` / ``Symmetries`` objects, ``Expr`` construction in every spelling
those files use (``×``, sites as tuples, spin arrow before or after the index), ``replace_indices``, ``adjoint``,
``+ - * scale`` with scalars on either side, ``==`` up to reordering.  (What needs a built basis -- ``basis.states``,
``index``, ``apply_*_to_basis_state``, ``eigsh`` -- is in the ``-m gpu`` tier: tests/test_gpu_parity.py.)"""
from __future__ import annotations

import numpy as np
import pytest

from lattice_symmetries_b200 import Expr, Symmetries, Symmetry
from lattice_symmetries_b200.expr import compile_terms

# python/test/test_api.py:45-47 (anisotropic 12-site kagome; the same string with "×" in python/run_tests.py:36-38)
KAGOME12 = (
    "1.0 σᶻ₀ σᶻ₁ + 1.0 σᶻ₀ σᶻ₃ + 1.0 σᶻ₀ σᶻ₈ + 1.0 σᶻ₀ σᶻ₁₀ + 2.0 σ⁺₀ σ⁻₁ + 2.0 σ⁺₀ σ⁻₃ + 2.0 σ⁺₀ σ⁻₈ + 2.0 σ⁺₀ σ⁻₁₀ + "
    "2.0 σ⁻₀ σ⁺₁ + 2.0 σ⁻₀ σ⁺₃ + 2.0 σ⁻₀ σ⁺₈ + 2.0 σ⁻₀ σ⁺₁₀ + 1.0 σᶻ₁ σᶻ₂ + 0.8 σᶻ₁ σᶻ₃ + 0.8 σᶻ₁ σᶻ₉ + 2.0 σ⁺₁ σ⁻₂ + "
    "1.6 σ⁺₁ σ⁻₃ + 1.6 σ⁺₁ σ⁻₉ + 2.0 σ⁻₁ σ⁺₂ + 1.6 σ⁻₁ σ⁺₃ + 1.6 σ⁻₁ σ⁺₉ + 1.0 σᶻ₂ σᶻ₄ + 1.0 σᶻ₂ σᶻ₉ + 1.0 σᶻ₂ σᶻ₁₀ + "
    "2.0 σ⁺₂ σ⁻₄ + 2.0 σ⁺₂ σ⁻₉ + 2.0 σ⁺₂ σ⁻₁₀ + 2.0 σ⁻₂ σ⁺₄ + 2.0 σ⁻₂ σ⁺₉ + 2.0 σ⁻₂ σ⁺₁₀ + 1.0 σᶻ₃ σᶻ₅ + 0.8 σᶻ₃ σᶻ₁₁ + "
    "2.0 σ⁺₃ σ⁻₅ + 1.6 σ⁺₃ σ⁻₁₁ + 2.0 σ⁻₃ σ⁺₅ + 1.6 σ⁻₃ σ⁺₁₁ + 0.8 σᶻ₄ σᶻ₆ + 1.0 σᶻ₄ σᶻ₇ + 0.8 σᶻ₄ σᶻ₁₀ + 1.6 σ⁺₄ σ⁻₆ + "
    "2.0 σ⁺₄ σ⁻₇ + 1.6 σ⁺₄ σ⁻₁₀ + 1.6 σ⁻₄ σ⁺₆ + 2.0 σ⁻₄ σ⁺₇ + 1.6 σ⁻₄ σ⁺₁₀ + 1.0 σᶻ₅ σᶻ₆ + 1.0 σᶻ₅ σᶻ₈ + 1.0 σᶻ₅ σᶻ₁₁ + "
    "2.0 σ⁺₅ σ⁻₆ + 2.0 σ⁺₅ σ⁻₈ + 2.0 σ⁺₅ σ⁻₁₁ + 2.0 σ⁻₅ σ⁺₆ + 2.0 σ⁻₅ σ⁺₈ + 2.0 σ⁻₅ σ⁺₁₁ + 1.0 σᶻ₆ σᶻ₇ + 0.8 σᶻ₆ σᶻ₈ + "
    "2.0 σ⁺₆ σ⁻₇ + 1.6 σ⁺₆ σ⁻₈ + 2.0 σ⁻₆ σ⁺₇ + 1.6 σ⁻₆ σ⁺₈ + 1.0 σᶻ₇ σᶻ₉ + 1.0 σᶻ₇ σᶻ₁₁ + 2.0 σ⁺₇ σ⁻₉ + 2.0 σ⁺₇ σ⁻₁₁ + "
    "2.0 σ⁻₇ σ⁺₉ + 2.0 σ⁻₇ σ⁺₁₁ + 0.8 σᶻ₈ σᶻ₁₀ + 1.6 σ⁺₈ σ⁻₁₀ + 1.6 σ⁻₈ σ⁺₁₀ + 0.8 σᶻ₉ σᶻ₁₁ + 1.6 σ⁺₉ σ⁻₁₁ + 1.6 σ⁻₉ σ⁺₁₁")


def _table(expr, n):
    return sorted((t.m, t.r, t.x, t.s, complex(np.round(t.v, 12))) for t in compile_terms(expr, n))


def test_symmetry():
    """python/test/test_api.py:19-29"""
    a = Symmetry([0, 1, 2], sector=0)
    assert a.sector == 0
    assert len(a) == 3
    assert a.permutation.tolist() == [0, 1, 2]
    assert a.json_object() == {"permutation": [0, 1, 2], "sector": 0}
    with pytest.raises(ValueError):
        Symmetry([0, 0, 1], sector=0)
    with pytest.raises(ValueError):
        Symmetry([1, 2, 0], sector=3)      # the periodicity of the permutation is 3: sectors 0, 1, 2


def test_symmetries():
    """python/test/test_api.py:32-35"""
    a = Symmetry([1, 2, 3, 0], sector=0)
    b = Symmetry([3, 2, 1, 0], sector=0)
    c = Symmetries([a, b])
    assert len(c) == 2 and [g.permutation.tolist() for g in c.generators] == [[1, 2, 3, 0], [3, 2, 1, 0]]
    assert len(c.elements) == 8                       # the dihedral group of the square
    with pytest.raises(ValueError):
        Symmetries([a, Symmetry([1, 0, 2], sector=0)])      # different numbers of sites
    with pytest.raises(ValueError):
        Symmetries([Symmetry([1, 2, 3, 0], sector=1), Symmetry([3, 2, 1, 0], sector=1)])   # incompatible sectors


def test_kagome_expression_is_invariant_under_its_symmetry():
    """python/test/test_api.py:44-51: ``expr == expr.replace_indices(...)`` for the right shift of the cluster."""
    expr = Expr(KAGOME12)
    right_shift = Symmetry([2, 10, 0, 4, 3, 7, 11, 5, 9, 8, 1, 6], sector=1)
    moved = expr.replace_indices(dict(zip(range(12), right_shift.permutation)))
    assert expr == moved and _table(expr, 12) == _table(moved, 12)
    assert Expr(KAGOME12.replace(" σ", " × σ", 1)) == expr          # the "×" spelling of python/run_tests.py:36
    wrong = expr.replace_indices(dict(zip(range(12), [1, 0] + list(range(2, 12)))))
    assert not (expr == wrong)
    Symmetries([right_shift])                                          # a valid one-generator group


def test_fermion_spellings_and_algebra():
    """python/run_tests.py:56-86 (issue01, create_hubbard_hamiltonian), python/test/test_api.py:84-96."""
    a = Expr("n↑₀ n↓₀", [(0,), (1,)])
    b = Expr("c↑₀ c†↑₁", [(0, 1)])
    c = Expr("c↓₀ c†↓₁", [(0, 1)])
    total = 2 * a + b + b.adjoint() + c + c.adjoint()
    same = Expr("2 n₀↑ n₀↓ + 2 n₁↑ n₁↓ + c₀↑ c†₁↑ + c₁↑ c†₀↑ + c₀↓ c†₁↓ + c₁↓ c†₀↓")
    assert total == same and _table(total, 2) == _table(same, 2)

    t, U = 1, 2
    operator = -t * Expr("c†↑₀ c↑₁", [(0, 1)])
    operator -= t * Expr("c†↑₁ c↑₀", [(0, 1)])
    operator -= t * Expr("c†↓₁ c↓₀", [(0, 1)])
    operator -= t * Expr("c†↓₀ c↓₁", [(0, 1)])
    operator += U * Expr("n↑₀ n↓₀", [(0,)])
    operator += U * Expr("n↑₁ n↓₁", [(1,)])
    one_line = Expr("- (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓) + 2.0 n₀↑ n₀↓ + 2.0 n₁↑ n₁↓")
    assert operator == one_line

    hopping = Expr("- (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓)")
    coulomb = Expr("4.0 n₀↑ n₀↓")
    expr = hopping + coulomb
    for i, j in [(1, 2), (2, 3), (3, 0)]:
        expr += hopping.replace_indices({0: i, 1: j})
    for i in [1, 2, 3]:
        expr += coulomb.replace_indices({0: i})
    ring = Expr("- (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓)", sites=[[0, 1], [1, 2], [2, 3], [3, 0]]) + \
        Expr("4.0 n₀↑ n₀↓", sites=[[0], [1], [2], [3]])
    assert expr == ring and _table(expr, 4) == _table(ring, 4)
    assert expr.scale(0.5) == 0.5 * expr == expr * 0.5


def test_scalar_sums_as_the_reference_writes_them():
    """python/test/test_api.py:9-16, 121-135: ``t1 * sum1(...) + t2 * sum1(...) + U * sum1(...)``."""
    def sum1(xs):
        s = None
        for x in xs:
            s = x if s is None else s + x
        return s

    nearest = [(0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3), (6, 7), (7, 8), (8, 6)]
    hopping = lambda i, j: Expr("c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓").replace_indices({0: i, 1: j})
    coulomb = lambda i: Expr("n₀↑ n₀↓").replace_indices({0: i})
    expr = -0.3251 * sum1(hopping(i, j) for i, j in nearest) + 2.8 * sum1(coulomb(i) for i in range(9))
    direct = Expr("-0.3251 (c†₁↑ c₀↑ + c†₀↑ c₁↑ + c†₁↓ c₀↓ + c†₀↓ c₁↓)", sites=nearest) + \
        Expr("2.8 n₀↑ n₀↓", sites=[[i] for i in range(9)])
    assert expr == direct
    terms = compile_terms(expr, 9)
    assert sum(1 for t in terms if t.x == 0) == 9 and sum(1 for t in terms if t.x != 0) == 4 * len(nearest)
