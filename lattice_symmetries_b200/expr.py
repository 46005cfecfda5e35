"""Operator expressions -> non-branching bitmask terms (host side).

The reference compiles expressions in Haskell (Parser.hs, Expr.hs:322-333,
Generator.hs:229-247, NonbranchingTerm.hs:24-58, Operator.hs:72-135) and hands
the hot path two ``ls_hs_nonbranching_terms`` tables (diagonal / off-diagonal,
kernels/lattice_symmetries_types.h:140-151).  GHC is unavailable, so this is a
small restatement of exactly that pipeline -- enough of the expression syntax
(``σˣ₀ σˣ₁``, ``2 (σ⁺₀ σ⁻₁ + σ⁺₁ σ⁻₀)``, ``c†₁↑ c₀↑``, ``4.0 n₀↑ n₀↓`` ...) for
the reference's own test and benchmark models.  It is *not* the symbolic
algebra engine (out of scope): products are expanded, each
product of generators is folded with the reference's composition rule and
equal (m, l, r, x, s) terms are merged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["Expr", "NonbranchingTerm", "compile_terms"]

_SUB = {c: str(i) for i, c in enumerate("₀₁₂₃₄₅₆₇₈₉")}
_SUP = {"ˣ": "x", "ʸ": "y", "ᶻ": "z", "⁺": "+", "⁻": "-"}


@dataclass(frozen=True)
class NonbranchingTerm:
    """(v, m, l, r, x, s) of NonbranchingTerm.hs:24-33: T|a> = v (-1)^{|a&s|}
    [a&m == r] |a^x>."""

    v: complex
    m: int
    l: int
    r: int
    x: int
    s: int

    def __matmul__(self, b: "NonbranchingTerm") -> "NonbranchingTerm":
        """Operator product self . b (NonbranchingTerm.hs:43-58)."""
        a = self
        m = a.m | b.m
        r = b.r | (a.r & ~b.m)
        l = a.l | (b.l & ~a.m)
        x = l ^ r
        s = (a.s ^ b.s) & ~m
        if (a.r ^ b.l) & a.m & b.m:
            v = 0j
        else:
            # b acts first: sign (-1)^{|r & s_b|}, then a sees r ^ x_b.  The
            # reference writes z = (r & s_b) ^ x_b, which agrees with this only
            # for the canonically ordered products its algebra layer always
            # produces (s_a within s_b); z = r ^ x_b is the operator identity for
            # any order, so unordered input needs no normal-ordering pass here.
            z = r ^ b.x
            p = bin((r & b.s) ^ (z & a.s)).count("1")
            v = (-1) ** p * a.v * b.v
        return NonbranchingTerm(v, m, l, r, x, s)


def _primitive(kind: str, op: str, bit: int) -> NonbranchingTerm:
    """Generator.hs:229-247."""
    b = 1 << bit
    if kind == "spin":
        if op == "z":
            return NonbranchingTerm(-1, 0, 0, 0, 0, b)
        if op == "+":
            return NonbranchingTerm(1, b, b, 0, b, 0)
        if op == "-":
            return NonbranchingTerm(1, b, 0, b, b, 0)
    else:
        if op == "n":
            return NonbranchingTerm(1, b, b, b, 0, 0)
        if op == "c†":
            return NonbranchingTerm(1, b, b, 0, b, b - 1)
        if op == "c":
            return NonbranchingTerm(1, b, 0, b, b, b - 1)
    raise ValueError(f"unknown generator {kind}:{op}")


# A generator is (kind, op, spin, site); spin is None | 0 (up) | 1 (down).
Gen = Tuple[str, str, Optional[int], int]
Monomial = Tuple[Gen, ...]


class Expr:
    """Polynomial in spin / fermion generators.  Mirrors the user-facing
    ``Expr`` of python/lattice_symmetries/__init__.py:466-583 (construction
    from a string with optional ``sites``, ``+ - *``, ``scale``,
    ``replace_indices``, ``adjoint``, ``==``)."""

    def __init__(self, expression="", sites: Optional[Sequence[Sequence[int]]] = None,
                 _terms: Optional[List[Tuple[complex, Monomial]]] = None):
        if _terms is not None:
            self.terms = _terms
            return
        base = _Parser(expression).parse()
        if sites is None:
            self.terms = base
        else:
            sites = [[int(i) for i in (row if np.ndim(row) else [row])] for row in sites]
            out: List[Tuple[complex, Monomial]] = []
            # Expr.hs:270-289 replicateSiteIndices: the distinct site indices of the
            # expression, in ascending order, are zipped with each row
            old = sorted({i for _, g in base for (_, _, _, i) in g})
            for row in sites:
                if len(row) != len(old):
                    raise ValueError(f"wrong number of site indices: {len(row)}; expected {len(old)}")
                mapping = dict(zip(old, row))
                out.extend(_replace(base, lambda s, i, mp=mapping: (s, mp[i])))
            self.terms = out

    # -- algebra ------------------------------------------------------------
    def __add__(self, other: "Expr") -> "Expr":
        return Expr(_terms=self.terms + other.terms)

    def __neg__(self) -> "Expr":
        return self.scale(-1)

    def __sub__(self, other: "Expr") -> "Expr":
        return self + (-other)

    def scale(self, c: complex) -> "Expr":
        return Expr(_terms=[(c * v, g) for v, g in self.terms])

    def __mul__(self, other):
        if isinstance(other, Expr):
            return Expr(_terms=[(va * vb, ga + gb) for va, ga in self.terms for vb, gb in other.terms])
        if np.isscalar(other):
            return self.scale(other)
        return NotImplemented

    def __rmul__(self, other):
        if np.isscalar(other):
            return self.scale(other)
        return NotImplemented

    def replace_indices(self, mapping: Dict) -> "Expr":
        """python/lattice_symmetries/__init__.py:500-549."""
        if not mapping:
            return self
        key = next(iter(mapping))
        if isinstance(key, (int, np.integer)):
            f = lambda s, i: (s, mapping.get(i, i))
        elif isinstance(key, str):
            sp = {"↑": 0, "↓": 1}
            f = lambda s, i: (sp[mapping[("↑", "↓")[s]]] if ("↑", "↓")[s] in mapping else s, i)
        else:
            sp = {"↑": 0, "↓": 1, 0: 0, 1: 1}
            mp = {(sp[a], b): (sp[c], d) for (a, b), (c, d) in mapping.items()}
            f = lambda s, i: mp.get((s, i), (s, i))
        return Expr(_terms=_replace(self.terms, f))

    def adjoint(self) -> "Expr":
        dag = {"+": "-", "-": "+", "z": "z", "n": "n", "c": "c†", "c†": "c"}
        return Expr(_terms=[
            (np.conj(v), tuple((k, dag[o], s, i) for (k, o, s, i) in reversed(g)))
            for v, g in self.terms
        ])

    def __eq__(self, other) -> bool:
        if not isinstance(other, Expr):
            return NotImplemented
        big = 1 + max([i for t in (self.terms + other.terms) for (_, _, _, i) in t[1]] + [0])
        a = _canonical(compile_terms(self, big))
        b = _canonical(compile_terms(other, big))
        keys = set(a) | set(b)
        return all(abs(a.get(k, 0) - b.get(k, 0)) < 1e-12 for k in keys)

    def __repr__(self):
        return f"Expr(<{len(self.terms)} monomials>)"


def _replace(terms, f):
    out = []
    for v, g in terms:
        new = []
        for (k, o, s, i) in g:
            s2, i2 = f(s, i)
            new.append((k, o, s2, i2))
        out.append((v, tuple(new)))
    return out


def _canonical(terms: Iterable[NonbranchingTerm]) -> Dict[tuple, complex]:
    """Fold sign bits that the (m, r) constraint fixes into v so that equal
    operators compare equal."""
    out: Dict[tuple, complex] = {}
    for t in terms:
        v = t.v * (-1) ** bin(t.r & t.s).count("1")
        key = (t.m, t.r, t.x, t.s & ~t.m)
        out[key] = out.get(key, 0) + v
    return {k: v for k, v in out.items() if abs(v) > 1e-14}


def compile_terms(expr: Expr, number_sites: int) -> List[NonbranchingTerm]:
    """Operator.hs:72-83 getNonbranchingTerms: flatten indices
    (Basis.hs:621-628: up -> k, down -> n + k), fold every product with ``<>``
    (Algebra.hs:132-135), merge equal bitmask signatures, drop zeros."""
    merged: Dict[tuple, complex] = {}
    order: List[tuple] = []
    for v, gens in expr.terms:
        if len(gens) == 0:
            t = NonbranchingTerm(v, 0, 0, 0, 0, 0)
        else:
            t = None
            for (kind, op, spin, site) in gens:
                bit = site if (spin is None or spin == 0) else number_sites + site
                p = _primitive(kind, op, bit)
                t = p if t is None else (t @ p)
            t = NonbranchingTerm(v * t.v, t.m, t.l, t.r, t.x, t.s)
        key = (t.m, t.l, t.r, t.x, t.s)
        if key not in merged:
            merged[key] = 0
            order.append(key)
        merged[key] += t.v
    return [NonbranchingTerm(complex(merged[k]), *k) for k in order if abs(merged[k]) > 1e-15]


# ----------------------------------------------------------------------------
class _Parser:
    """Recursive-descent parser for the subset of Parser.hs used by the
    reference's tests, examples and YAML models."""

    def __init__(self, text: str):
        self.s = text
        self.i = 0

    def _ws(self):
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1

    def _peek(self) -> str:
        self._ws()
        return self.s[self.i] if self.i < len(self.s) else ""

    def parse(self):
        out = self._sum()
        self._ws()
        if self.i != len(self.s):
            raise ValueError(f"unexpected {self.s[self.i]!r} at {self.i} in {self.s!r}")
        return out

    def _sum(self):
        terms: List[Tuple[complex, Monomial]] = []
        first = True
        while True:
            c = self._peek()
            if c == "" or c == ")":
                break
            sign = 1
            if c in "+-":
                self.i += 1
                sign = -1 if c == "-" else 1
            elif not first:
                break
            terms.extend((sign * v, g) for v, g in self._product())
            first = False
        return terms

    def _number(self) -> Optional[complex]:
        self._ws()
        j = self.i
        while j < len(self.s) and (self.s[j].isdigit() or self.s[j] in ".eE" or
                                   (self.s[j] in "+-" and j > self.i and self.s[j - 1] in "eE")):
            j += 1
        if j == self.i:
            return None
        val = float(self.s[self.i:j])
        self.i = j
        if self.s.startswith("im", self.i):
            self.i += 2
            return complex(0, val)
        if self.i < len(self.s) and self.s[self.i] in "Iⅈ":
            self.i += 1
            return complex(0, val)
        return complex(val, 0)

    def _complex_in_parens(self) -> Optional[complex]:
        """``(a + bim)`` literal."""
        save = self.i
        self._ws()
        if self._peek() != "(":
            return None
        self.i += 1
        a = self._number()
        if a is None:
            self.i = save
            return None
        c = self._peek()
        if c in "+-":
            self.i += 1
            b = self._number()
            if b is None or b.real != 0 or self._peek() != ")":
                self.i = save
                return None
            self.i += 1
            return a + (b if c == "+" else -b)
        if c == ")":
            self.i += 1
            return a
        self.i = save
        return None

    def _product(self):
        coeff = self._number()
        if coeff is None:
            coeff = self._complex_in_parens()
        if coeff is None:
            coeff = 1
        elif self._peek() == "×":
            self.i += 1
        acc: List[Tuple[complex, Monomial]] = [(coeff, ())]
        n_factors = 0
        while True:
            c = self._peek()
            if c == "(":
                self.i += 1
                inner = self._sum()
                if self._peek() != ")":
                    raise ValueError(f"expected ')' at {self.i} in {self.s!r}")
                self.i += 1
                acc = [(va * vb, ga + gb) for va, ga in acc for vb, gb in inner]
            elif c == "×":
                self.i += 1
                continue
            elif c in ("σ", "S", "c", "n") or self.s.startswith("\\sigma", self.i):
                prim = self._primitive()
                acc = [(va * vb, ga + gb) for va, ga in acc for vb, gb in prim]
            else:
                break
            n_factors += 1
        return acc

    def _subscript(self) -> int:
        j = self.i
        digits = ""
        if j < len(self.s) and self.s[j] == "_":
            j += 1
        while j < len(self.s) and (self.s[j] in _SUB or self.s[j].isdigit()):
            digits += _SUB.get(self.s[j], self.s[j])
            j += 1
        if not digits:
            raise ValueError(f"expected a site index at {self.i} in {self.s!r}")
        self.i = j
        return int(digits)

    def _primitive(self):
        self._ws()
        if self.s.startswith("\\sigma", self.i):
            self.i += 6
            c = "σ"
        else:
            c = self.s[self.i]
            self.i += 1
        if c in ("σ", "S"):
            ch = self.s[self.i]
            if ch == "^":
                self.i += 1
                ch = self.s[self.i]
            op = _SUP.get(ch, ch)
            if op not in ("x", "y", "z", "+", "-"):
                raise ValueError(f"invalid spin operator at {self.i} in {self.s!r}")
            self.i += 1
            site = self._subscript()
            scale = 0.5 if c == "S" else 1.0  # Expr.hs:332-333
            g = lambda o: (("spin", o, None, site),)
            if op == "x":  # Expr.hs:327
                return [(scale, g("+")), (scale, g("-"))]
            if op == "y":  # Expr.hs:328-331: -i (s+ - s-)
                return [(-1j * scale, g("+")), (1j * scale, g("-"))]
            return [(scale, g(op))]
        # fermions
        if c == "n":
            op = "n"
        else:
            if self.i < len(self.s) and self.s[self.i] == "†":
                self.i += 1
                op = "c†"
            else:
                op = "c"
        # both orders occur in the reference: "c†₁↑" (test/04_hubbard_square/hamiltonian.yaml:7)
        # and "c†↑₀" (python/run_tests.py:80-86)
        spin = None
        if self.i < len(self.s) and self.s[self.i] in "↑↓":
            spin = 0 if self.s[self.i] == "↑" else 1
            self.i += 1
        site = self._subscript()
        if spin is None and self.i < len(self.s) and self.s[self.i] in "↑↓":
            spin = 0 if self.s[self.i] == "↑" else 1
            self.i += 1
        return [(1.0, (("fermion", op, spin, site),))]
