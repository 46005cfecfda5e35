"""Host-side symmetry groups: closure, characters and Benes networks.

Mirrors the *interface* of the reference host for this path -- ``Symmetry`` /
``Symmetries`` of python/lattice_symmetries/__init__.py:66-131, whose payloads
are built by the Haskell host (haskell/src/LatticeSymmetries/Group.hs:58-208,
Benes.hs:299-377).  GHC is not available, so this module produces the very
``ls_hs_permutation_group`` tables (kernels/lattice_symmetries_types.h:109-119)
the Haskell host would hand to ``ls_internal_create_halide_kernel_data``:

* group elements sorted ascending by (permutation, phase)   (Group.hs:174-183)
* characters cos/sin(-2 pi phi)                              (Group.hs:115-116)
* one Benes network per element with the shared shift list
  [1, 2, ..., n_pad/2, ..., 2, 1], masks laid out [depth][number_masks]
  (Benes.hs:299-301, 356-377); n_pad = next power of two (Benes.hs:308-314).

The network solver below is our own (recursive 2-colouring); any valid network
is acceptable because only g.x is observable (permuteBits' semantics,
Benes.hs:338-345: output bit i = input bit p[i]).
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Iterable, List, Sequence, Tuple

import numpy as np

__all__ = ["Symmetry", "Symmetries", "benes_network", "apply_benes", "permute_bits_naive"]


def _check_permutation(p: Sequence[int]) -> Tuple[int, ...]:
    p = tuple(int(i) for i in p)
    if len(p) == 0 or sorted(p) != list(range(len(p))):
        raise ValueError(f"invalid permutation: {list(p)}")  # Benes.hs:96-101
    return p


def _periodicity(p: Tuple[int, ...]) -> int:
    """Group.hs:47-53 getPeriodicity."""
    identity = tuple(range(len(p)))
    n, q = 1, p
    while q != identity:
        q = tuple(q[i] for i in p)
        n += 1
    return n


def _compose(a: Tuple[int, ...], b: Tuple[int, ...]) -> Tuple[int, ...]:
    """Benes.hs:103-104: (x <> ys)[i] = ys[x[i]]."""
    return tuple(b[i] for i in a)


class Symmetry:
    """A lattice symmetry: a site permutation plus a sector (Group.hs:58-96;
    python/lattice_symmetries/__init__.py:66-105)."""

    def __init__(self, permutation: Iterable[int], sector: int, _phase: Fraction | None = None):
        self._perm = _check_permutation(list(permutation))
        if _phase is None:
            periodicity = _periodicity(self._perm)
            sector = int(sector)
            if sector < 0 or sector >= periodicity:
                raise ValueError(
                    f"invalid sector: {sector}; permutation has periodicity {periodicity}"
                )
            self._phase = Fraction(sector, periodicity)
        else:
            self._phase = _phase

    @property
    def permutation(self) -> np.ndarray:
        return np.array(self._perm, dtype=np.int32)

    @property
    def phase(self) -> Fraction:
        return self._phase

    @property
    def sector(self) -> int:
        s = self._phase * _periodicity(self._perm)
        if s.denominator != 1:
            raise ValueError("invalid symmetry")
        return int(s.numerator)

    def __len__(self) -> int:
        return len(self._perm)

    def _key(self):
        return (self._perm, self._phase)

    def __mul__(self, other: "Symmetry") -> "Symmetry":
        """Group.hs:95-96: (pa, la) <> (pb, lb) = (pa <> pb, (la + lb) mod 1)."""
        phase = self._phase + other._phase
        if phase >= 1:
            phase -= phase.numerator // phase.denominator
        return Symmetry(_compose(self._perm, other._perm), 0, _phase=phase)

    def __eq__(self, other):
        return isinstance(other, Symmetry) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def json_object(self):
        return {"permutation": list(self._perm), "sector": self.sector}

    def __repr__(self):
        return f"Symmetry({list(self._perm)}, sector={self.sector})"


# ----------------------------------------------------------------------------
# Benes networks
# ----------------------------------------------------------------------------
def _next_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def benes_shifts(number_bits: int) -> List[int]:
    """[1, 2, ..., n_pad/2, ..., 2, 1] (Benes.hs:299-301)."""
    n = _next_pow2(number_bits)
    up = []
    d = 1
    while 2 * d <= n:
        up.append(d)
        d *= 2
    return up + up[-2::-1]


def benes_network(perm: Sequence[int]) -> Tuple[List[int], List[int]]:
    """Return ``(masks, shifts)`` such that applying ``bit_permute_step`` with
    ``(masks[k], shifts[k])`` for k = 0.. maps x to y with y_i = x_{perm[i]}."""
    p = list(perm)
    n = _next_pow2(len(p))
    p = p + list(range(len(p), n))  # extendToPowerOfTwo, Benes.hs:308-314
    shifts = benes_shifts(n)
    depth = len(shifts)
    masks = [0] * depth
    if depth == 0:
        return masks, shifts
    levels = (depth + 1) // 2  # log2(n)

    # (level, base, q): sub-problem on positions base + d*k, d = 2**level,
    # with out[k] = in[q[k]].
    stack = [(0, 0, p)]
    while stack:
        level, base, q = stack.pop()
        d = 1 << level
        m = len(q)
        if m == 2:
            if q[0] == 1:
                masks[levels - 1] |= 1 << base
            continue
        outpos = [0] * m
        for k, s in enumerate(q):
            outpos[s] = k
        color = [-1] * m
        for s0 in range(m):
            if color[s0] != -1:
                continue
            s = s0
            while color[s] == -1:
                color[s] = 0
                partner = s ^ 1  # shares the input switch -> other half
                color[partner] = 1
                s = q[outpos[partner] ^ 1]  # shares partner's output switch
        q_sub = ([0] * (m // 2), [0] * (m // 2))
        for a in range(m // 2):
            if color[2 * a] == 1:  # input switch a crossed
                masks[level] |= 1 << (base + d * 2 * a)
        for b in range(m // 2):
            if color[q[2 * b]] == 1:  # output switch b crossed
                masks[depth - 1 - level] |= 1 << (base + d * 2 * b)
            for k in (2 * b, 2 * b + 1):
                s = q[k]
                q_sub[color[s]][b] = s // 2
        stack.append((level + 1, base, q_sub[0]))
        stack.append((level + 1, base + d, q_sub[1]))
    return masks, shifts


def apply_benes(masks: Sequence[int], shifts: Sequence[int], x: int) -> int:
    """Benes.hs:325-336 permuteBits (bitPermuteStep folded over the stages)."""
    for m, d in zip(masks, shifts):
        y = ((x >> d) ^ x) & m
        x = (x ^ y) ^ (y << d)
    return x


def permute_bits_naive(perm: Sequence[int], x: int) -> int:
    """Benes.hs:338-345 permuteBits': output bit i = input bit perm[i]."""
    y = 0
    for i, pi in enumerate(perm):
        if (x >> pi) & 1:
            y |= 1 << i
    return y


# ----------------------------------------------------------------------------
class Symmetries:
    """Closure of a list of generators (Group.hs:106-132, 174-183;
    python/lattice_symmetries/__init__.py:108-131)."""

    def __init__(self, generators: Sequence[Symmetry] = ()):
        self._generators = list(generators)
        self.elements: List[Symmetry] = []
        if self._generators:
            n = len(self._generators[0])
            if any(len(g) != n for g in self._generators):
                raise ValueError("symmetries have different number of sites")
            self.elements = self._closure(n)
            self._check_consistency()

    def _closure(self, n: int) -> List[Symmetry]:
        identity = Symmetry(range(n), 0)
        interior: set = set()
        boundary = {identity}
        while boundary:
            interior |= boundary
            boundary = {h * g for h in boundary for g in self._generators} - interior
        return sorted(interior, key=lambda s: s._key())  # Set.toAscList

    def _check_consistency(self) -> None:
        members = set(self.elements)
        for a in self.elements:
            for b in self.elements:
                c = a * b
                ok = c in members and (c.phase * _periodicity(c._perm)).denominator == 1
                if not ok:
                    raise ValueError("incompatible symmetries")

    def __len__(self) -> int:
        return len(self._generators)

    @property
    def generators(self) -> List[Symmetry]:
        return self._generators

    @property
    def number_bits(self) -> int:
        """Group.hs symmetriesGetNumberBits: 0 for the empty group."""
        return len(self.elements[0]) if self.elements else 0

    @property
    def is_empty(self) -> bool:
        return not self.elements

    def json_object(self):
        return [g.json_object() for g in self._generators]

    def permutations(self) -> np.ndarray:
        return np.array([s._perm for s in self.elements], dtype=np.int32).reshape(
            len(self.elements), self.number_bits
        )

    def characters(self) -> Tuple[np.ndarray, np.ndarray]:
        """Group.hs:115-116: cos / sin of (-2 pi phi)."""
        re = np.array([math.cos(-2 * math.pi * float(s.phase)) for s in self.elements])
        im = np.array([math.sin(-2 * math.pi * float(s.phase)) for s in self.elements])
        return re.astype(np.float64), im.astype(np.float64)

    def is_real(self) -> bool:
        return bool(np.all(self.characters()[1] == 0))

    def tables(self):
        """The ``ls_hs_permutation_group`` payload: (number_bits, shifts u64[depth],
        masks u64[depth][number_masks], eigvals_re, eigvals_im)."""
        nbits = self.number_bits
        G = len(self.elements)
        shifts = benes_shifts(nbits) if G else []
        depth = len(shifts)
        masks = np.zeros((depth, G), dtype=np.uint64)
        for j, s in enumerate(self.elements):
            ms, _ = benes_network(s._perm)
            for k in range(depth):
                masks[k, j] = ms[k]
        re, im = self.characters() if G else (np.zeros(0), np.zeros(0))
        return (
            nbits,
            np.array(shifts, dtype=np.uint64),
            np.ascontiguousarray(masks),
            np.ascontiguousarray(re, dtype=np.float64),
            np.ascontiguousarray(im, dtype=np.float64),
        )
