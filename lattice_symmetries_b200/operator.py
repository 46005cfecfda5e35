"""Quantum operators: host-side mirror of the reference's ``Operator``
(python/lattice_symmetries/__init__.py:591-752).

The reference's Haskell host compiles the expression into two
``ls_hs_nonbranching_terms`` tables (haskell/src/LatticeSymmetries/Operator.hs:
109-155: diagonal terms have x == 0) and fills ``ls_hs_operator``; this module
does the same with :mod:`expr` and calls the C ABI for everything else.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Tuple

import numpy as np

from . import _lib
from ._lib import lib
from .basis import Basis
from .expr import Expr, compile_terms

try:  # scipy is optional; the reference derives from LinearOperator so eigsh works
    from scipy.sparse.linalg import LinearOperator as _Base
except Exception:  # pragma: no cover
    _Base = object

__all__ = ["Operator"]


def _terms_struct(terms, number_bits: int, keep: list):
    """Operator.hs:109-135 createCnonbranching_terms (NULL when empty)."""
    if not terms:
        return None
    v = np.array([[t.v.real, t.v.imag] for t in terms], dtype=np.float64)
    cols = {k: np.array([getattr(t, k) for t in terms], dtype=np.uint64) for k in "mlrxs"}
    s = _lib.ls_hs_nonbranching_terms()
    s.number_terms = len(terms)
    s.number_bits = number_bits
    s.v = v.ctypes.data
    for k, arr in cols.items():
        setattr(s, k, arr.ctypes.data)
    keep.extend([v, cols, s])
    return C.pointer(s)


def _external_to_numpy(arr: _lib.chpl_external_array, dtype) -> np.ndarray:
    """Copy out and release (python/lattice_symmetries/__init__.py:419-424)."""
    dtype = np.dtype(dtype)
    n = int(arr.num_elts)
    if n == 0 or not arr.elts:
        out = np.zeros(0, dtype=dtype)
    else:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(arr.elts)
        out = np.frombuffer(buf, dtype=dtype).copy()
    lib.ls_hs_internal_destroy_external_array(C.byref(arr))
    return out


def _release_operator(address: int) -> None:
    try:
        lib.ls_b200_operator_release(C.c_void_p(address))
    except Exception:  # interpreter shutdown
        pass


class Operator(_Base):
    def __init__(self, basis: Basis, expression: Expr):
        if not isinstance(basis, Basis):
            raise TypeError(f"expected a 'Basis', but got '{type(basis)}'")
        if not isinstance(expression, Expr):
            raise TypeError(f"expected a 'Expr', but got '{type(expression)}'")
        self._basis = basis
        self._expression = expression
        terms = compile_terms(expression, basis.number_sites)
        self._diag_terms = [t for t in terms if t.x == 0]       # NonbranchingTerm.hs:71-72
        self._off_diag_terms = [t for t in terms if t.x != 0]
        self._keep: list = []
        op = _lib.ls_hs_operator()
        op.refcount = 1
        op.basis = C.pointer(basis._payload)
        off = _terms_struct(self._off_diag_terms, basis.number_bits, self._keep)
        diag = _terms_struct(self._diag_terms, basis.number_bits, self._keep)
        if off is not None:
            op.off_diag_terms = off
        if diag is not None:
            op.diag_terms = diag
        self._payload = op
        # the library caches device copies of the term tables keyed by the struct's address: drop them with the object
        # (a later operator may be allocated at the same address)
        weakref.finalize(self, _release_operator, C.addressof(op))

    @property
    def basis(self) -> Basis:
        return self._basis

    @property
    def expression(self) -> Expr:
        return self._expression

    @property
    def number_off_diag_terms(self) -> int:
        return len(self._off_diag_terms)

    @property
    def number_diag_terms(self) -> int:
        return len(self._diag_terms)

    @property
    def max_number_off_diag(self) -> int:
        """Operator.hs:190-193: number of distinct x masks."""
        return len({t.x for t in self._off_diag_terms})

    # -- algebra (delegates to Expr like the reference) ---------------------------
    def __add__(self, other):
        return Operator(self._basis, self._expression + other.expression)

    def __sub__(self, other):
        return Operator(self._basis, self._expression - other.expression)

    def scale(self, coeff: complex) -> "Operator":
        return Operator(self._basis, self._expression.scale(coeff))

    def __mul__(self, other):
        if isinstance(other, Operator):
            return Operator(self._basis, self._expression * other.expression)
        return NotImplemented

    def __rmul__(self, other):
        if np.isscalar(other):
            return self.scale(other)
        return NotImplemented

    def __matmul__(self, other):
        return self.apply_to_state_vector(other)

    def __repr__(self):
        return "<Operator defined on {}>".format(self._basis.__class__.__name__)

    # -- rows of H ----------------------------------------------------------------
    def apply_diag_to_basis_state(self, state: int) -> float:
        arr = (C.c_uint64 * 1)(int(state))
        coeffs = _lib.chpl_external_array()
        kernels = lib.ls_hs_internal_get_chpl_kernels()
        kernels.contents.operator_apply_diag(C.byref(self._payload), 1, arr, C.byref(coeffs), 0)
        _lib.check_error()
        return float(_external_to_numpy(coeffs, np.float64)[0])

    def apply_off_diag_to_basis_state(self, state: int) -> List[Tuple[complex, int]]:
        arr = (C.c_uint64 * 1)(int(state))
        betas, coeffs, offsets = (_lib.chpl_external_array() for _ in range(3))
        kernels = lib.ls_hs_internal_get_chpl_kernels()
        kernels.contents.operator_apply_off_diag(
            C.byref(self._payload), 1, arr, C.byref(betas), C.byref(coeffs), C.byref(offsets), 0)
        _lib.check_error()
        offsets_arr = _external_to_numpy(offsets, np.int64)
        n = int(offsets_arr[1])
        betas_arr = _external_to_numpy(betas, np.uint64)[:n]
        coeffs_arr = _external_to_numpy(coeffs, np.complex128)[:n]
        return [(complex(c), int(b)) for c, b in zip(coeffs_arr, betas_arr)]

    def apply_off_diag(self, states, xs=None):
        """Batched ls_internal_operator_apply_off_diag_x1 (kernels/reference.c:97-134):
        returns (betas, coeffs, offsets)."""
        states = np.ascontiguousarray(states, dtype=np.uint64)
        n = states.shape[0]
        T = max(self.number_off_diag_terms, 1)
        betas = np.zeros(n * T, dtype=np.uint64)
        coeffs = np.zeros(n * T, dtype=np.complex128)
        offsets = np.zeros(n + 1, dtype=np.int64)
        xs_p = None
        if xs is not None:
            xs = np.ascontiguousarray(xs, dtype=np.float64)
            xs_p = xs.ctypes.data_as(_lib.f64_p)
        lib.ls_internal_operator_apply_off_diag_x1(
            C.byref(self._payload), n, states.ctypes.data_as(_lib.u64_p), betas.ctypes.data_as(_lib.u64_p),
            coeffs.ctypes.data, offsets.ctypes.data_as(_lib.i64_p), xs_p)
        _lib.check_error()
        total = int(offsets[-1])
        return betas[:total], coeffs[:total], offsets

    def apply_off_diag_projected(self, states, with_indices: bool = False):
        """Rows of the symmetry-projected operator, batched on the device (extension; the reference
        computes them inside its matvec only, chapel/src/BatchedOperator.chpl:207-253): for every ``states[i]``
        the representatives of its images and the matrix elements  chi c n(beta) / n(alpha_i).
        Returns (representatives, coefficients, offsets[, indices])."""
        states = np.ascontiguousarray(states, dtype=np.uint64)
        count = states.shape[0]
        cap = max(1, count * max(self.number_off_diag_terms, 1))
        reps = np.zeros(cap, dtype=np.uint64)
        coeffs = np.zeros(cap, dtype=np.complex128)
        offsets = np.zeros(count + 1, dtype=np.int64)
        indices = np.zeros(cap, dtype=np.int64) if with_indices else None
        if with_indices:
            self._check_basis_is_built("apply_off_diag_projected")
        total = lib.ls_b200_operator_apply_off_diag_projected(
            C.byref(self._payload), count, states.ctypes.data_as(_lib.u64_p), reps.ctypes.data_as(_lib.u64_p),
            coeffs.ctypes.data, offsets.ctypes.data_as(_lib.i64_p),
            indices.ctypes.data_as(_lib.i64_p) if with_indices else None)
        _lib.check_error()
        if total < 0:
            raise RuntimeError("ls_b200_operator_apply_off_diag_projected failed")
        out = (reps[:total], coeffs[:total], offsets)
        return out + (indices[:total],) if with_indices else out

    def apply_diag(self, states, xs=None) -> np.ndarray:
        """Batched ls_internal_operator_apply_diag_x1 (kernels/reference.c:67-95)."""
        states = np.ascontiguousarray(states, dtype=np.uint64)
        ys = np.zeros(states.shape[0], dtype=np.float64)
        xs_p = None
        if xs is not None:
            xs = np.ascontiguousarray(xs, dtype=np.float64)
            xs_p = xs.ctypes.data_as(_lib.f64_p)
        lib.ls_internal_operator_apply_diag_x1(
            C.byref(self._payload), states.shape[0], states.ctypes.data_as(_lib.u64_p),
            ys.ctypes.data_as(_lib.f64_p), xs_p)
        _lib.check_error()
        return ys

    # -- matvec --------------------------------------------------------------------
    def _check_basis_is_built(self, attribute):
        if not self._basis.is_built:
            raise AttributeError(
                "'Operator' object has no attribute '{}' (did you forget to build the basis?)".format(attribute))

    def apply_to_state_vector(self, vector: np.ndarray) -> np.ndarray:
        """y = H x through the registered ``matrix_vector_product`` kernel with
        host buffers (python/lattice_symmetries/__init__.py:715-729)."""
        self._check_basis_is_built("apply_to_state_vector")
        if vector.dtype != np.float64:
            raise TypeError(
                "expected a NDArray[float64], but got {}[{}]".format(type(vector).__name__, type(vector.dtype)))
        vector = np.ascontiguousarray(vector)
        dim = self._basis.number_states
        # extension: a (k, dim) block of vectors shares one pass over the matrix elements (the reference halts
        # on numVectors != 1, chapel/src/DistributedMatrixVector.chpl:1096-1097)
        if vector.shape != (dim,) and not (vector.ndim == 2 and vector.shape[1] == dim and vector.shape[0] >= 1):
            raise ValueError(f"expected a vector of shape ({dim},), got {vector.shape}")
        number_vectors = 1 if vector.ndim == 1 else vector.shape[0]
        out = np.empty_like(vector)
        kernels = lib.ls_hs_internal_get_chpl_kernels()
        kernels.contents.matrix_vector_product(
            C.byref(self._payload), number_vectors, vector.ctypes.data_as(_lib.f64_p), out.ctypes.data_as(_lib.f64_p))
        _lib.check_error()
        return out

    def matvec_device_phase(self, phase: int, x_ptr: int = 0, y_ptr: int = 0, row_begin: int = 0, row_end: int = -1,
                            complex_vectors: bool = False) -> None:
        """Phase 1: canonicalise the matrix elements of the row range (independent of x); phase 2: rank, gather, sum."""
        self._check_basis_is_built("matvec_device_phase")
        if row_end < 0:
            row_end = self._basis.number_states
        status = lib.ls_b200_matvec_device_phase(
            C.byref(self._payload), int(row_begin), int(row_end), x_ptr or None, y_ptr or None,
            1 if complex_vectors else 0, int(phase))
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_matvec_device_phase failed")

    def matvec_block_device(self, number_vectors: int, x_ptr: int, x_stride: int, y_ptr: int, y_stride: int,
                            row_begin: int = 0, row_end: int = -1, sync: bool = False) -> None:
        """Block matvec on device-resident float64 vectors: vector v is x_ptr + 8 v x_stride -> y_ptr + 8 v y_stride;
        all vectors share one canonicalisation + ranking pass."""
        self._check_basis_is_built("matvec_block_device")
        if row_end < 0:
            row_end = self._basis.number_states
        status = lib.ls_b200_matvec_block_device(
            C.byref(self._payload), int(row_begin), int(row_end), int(number_vectors), x_ptr, int(x_stride), y_ptr,
            int(y_stride))
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_matvec_block_device failed")
        if sync:
            lib.ls_b200_matvec_sync()
            _lib.check_error()

    def matvec_device(self, x_ptr: int, y_ptr: int, row_begin: int = 0, row_end: int = -1,
                      complex_vectors: bool = False, sync: bool = False) -> None:
        """y[row_begin:row_end] = (H x)[row_begin:row_end] with x (length dim) and y
        in DEVICE memory (raw pointers, e.g. ``tensor.data_ptr()``); asynchronous on
        the library stream unless ``sync``."""
        self._check_basis_is_built("matvec_device")
        if row_end < 0:
            row_end = self._basis.number_states
        fn = lib.ls_b200_matvec_device_c128 if complex_vectors else lib.ls_b200_matvec_device
        status = fn(C.byref(self._payload), int(row_begin), int(row_end), x_ptr, y_ptr)
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_matvec_device failed")
        if sync:
            lib.ls_b200_matvec_sync()
            _lib.check_error()

    def count_matrix_elements(self, row_begin: int = 0, row_end: int = -1) -> int:
        """Off-diagonal matrix elements in a row range (the unit of the matvec metric)."""
        self._check_basis_is_built("count_matrix_elements")
        if row_end < 0:
            row_end = self._basis.number_states
        n = int(lib.ls_b200_count_matrix_elements(C.byref(self._payload), int(row_begin), int(row_end)))
        _lib.check_error()
        return n

    @property
    def _xp(self):
        """Array namespace newer scipy (>= 1.17) expects LinearOperator.__init__ to have
        set; like the reference class we never call it (shape is only known after build)."""
        try:
            from scipy._lib._array_api import np_compat
        except ImportError:  # older scipy never asks
            return np
        return np_compat

    @property
    def dtype(self):
        self._check_basis_is_built("dtype")
        return np.dtype("float64")

    @property
    def shape(self):
        self._check_basis_is_built("shape")
        n = self._basis.number_states
        return (n, n)

    def _matvec(self, x):
        return self.apply_to_state_vector(np.ascontiguousarray(x, dtype=np.float64).reshape(-1))
