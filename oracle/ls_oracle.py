"""ctypes front-end of the CPU oracle (oracle/ls_oracle.c -> liboracle.so) and of
the compiled reference sources (oracle/_ref/libref.so = the reference's own
kernels/indexing.c + kernels/reference.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(lattice_symmetries_b200/) never imports this module.

Inputs are plain numpy tables: the ``ls_hs_permutation_group`` payload
(masks u64[depth][|G|], shifts u64[depth], eigvals re/im) and the
``ls_hs_nonbranching_terms`` columns (v, m, l, r, x, s).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
LIBORACLE = HERE / "liboracle.so"
LIBREF = HERE / "_ref" / "libref.so"

u64_p = C.POINTER(C.c_uint64)
f64_p = C.POINTER(C.c_double)
i64_p = C.POINTER(C.c_int64)
u8_p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/libref.so when /root/reference is present)."""
    if force or not LIBORACLE.exists() or LIBORACLE.stat().st_mtime < (HERE / "ls_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)
    if Path("/root/reference/kernels").is_dir() and (force or not LIBREF.exists()):
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)


class oracle_group(C.Structure):
    _fields_ = [
        ("number_bits", C.c_int), ("depth", C.c_int), ("number_masks", C.c_int), ("spin_inversion", C.c_int),
        ("masks", C.c_void_p), ("shifts", C.c_void_p), ("eigvals_re", C.c_void_p), ("eigvals_im", C.c_void_p),
    ]


class oracle_basis(C.Structure):
    _fields_ = [
        ("number_sites", C.c_int), ("number_particles", C.c_int), ("number_up", C.c_int),
        ("particle_type", C.c_int), ("spin_inversion", C.c_int), ("has_permutation_symmetries", C.c_int),
        ("group", oracle_group),
    ]


class oracle_terms(C.Structure):
    _fields_ = [
        ("number_terms", C.c_int), ("v", C.c_void_p), ("m", C.c_void_p), ("l", C.c_void_p), ("r", C.c_void_p),
        ("x", C.c_void_p), ("s", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIBORACLE))
        L.oracle_state_info.argtypes = [C.POINTER(oracle_group), C.c_ssize_t, u64_p, u64_p, f64_p, f64_p]
        L.oracle_is_representative.argtypes = [C.POINTER(oracle_group), C.c_ssize_t, u64_p, u8_p, f64_p]
        L.oracle_fixed_hamming_state_to_index.argtypes = [C.c_uint64]
        L.oracle_fixed_hamming_state_to_index.restype = C.c_int64
        L.oracle_fixed_hamming_index_to_state.argtypes = [C.c_int64, C.c_int]
        L.oracle_fixed_hamming_index_to_state.restype = C.c_uint64
        L.oracle_min_state_estimate.argtypes = [C.POINTER(oracle_basis)]
        L.oracle_min_state_estimate.restype = C.c_uint64
        L.oracle_max_state_estimate.argtypes = [C.POINTER(oracle_basis)]
        L.oracle_max_state_estimate.restype = C.c_uint64
        L.oracle_enumerate_states.argtypes = [C.POINTER(oracle_basis), C.POINTER(C.c_uint64)]
        L.oracle_enumerate_states.restype = C.c_void_p
        L.oracle_enumerate_range.argtypes = [C.POINTER(oracle_basis), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_enumerate_range.restype = C.c_void_p
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_index_create.argtypes = [u64_p, C.c_ssize_t, C.c_int, C.c_int]
        L.oracle_index_create.restype = C.c_void_p
        L.oracle_index_destroy.argtypes = [C.c_void_p]
        L.oracle_state_index.argtypes = [C.c_void_p, C.c_ssize_t, u64_p, i64_p]
        L.oracle_apply_diag.argtypes = [C.POINTER(oracle_terms), C.c_ssize_t, u64_p, f64_p, f64_p]
        L.oracle_apply_off_diag.argtypes = [C.POINTER(oracle_terms), C.c_ssize_t, u64_p, u64_p, f64_p, i64_p, f64_p]
        L.oracle_matvec.argtypes = [
            C.POINTER(oracle_basis), C.POINTER(oracle_terms), C.POINTER(oracle_terms), C.c_void_p, u64_p,
            C.c_ssize_t, C.c_ssize_t, C.c_ssize_t, f64_p, f64_p, C.c_int]
        L.oracle_matvec.restype = C.c_int64
        L.oracle_matvec_strided.argtypes = [
            C.POINTER(oracle_basis), C.POINTER(oracle_terms), C.POINTER(oracle_terms), C.c_void_p, u64_p,
            C.c_ssize_t, C.c_ssize_t, C.c_ssize_t, C.c_ssize_t, f64_p, f64_p, C.c_int]
        L.oracle_matvec_strided.restype = C.c_int64
        L.oracle_set_reference_hooks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class Group:
    """ls_hs_permutation_group payload + spin inversion."""

    def __init__(self, number_bits: int, shifts, masks, re, im, spin_inversion: int = 0):
        self.shifts = np.ascontiguousarray(shifts, dtype=np.uint64)
        self.masks = np.ascontiguousarray(masks, dtype=np.uint64)
        self.re = np.ascontiguousarray(re, dtype=np.float64)
        self.im = np.ascontiguousarray(im, dtype=np.float64)
        g = oracle_group()
        g.number_bits = int(number_bits)
        g.depth = len(self.shifts)
        g.number_masks = len(self.re)
        g.spin_inversion = int(spin_inversion or 0)
        g.masks = self.masks.ctypes.data if self.masks.size else None
        g.shifts = self.shifts.ctypes.data if self.shifts.size else None
        g.eigvals_re = self.re.ctypes.data if self.re.size else None
        g.eigvals_im = self.im.ctypes.data if self.im.size else None
        self.c = g

    @classmethod
    def from_symmetries(cls, symmetries, number_sites: int, spin_inversion: Optional[int]):
        nbits, shifts, masks, re, im = symmetries.tables()
        return cls(nbits if len(re) else number_sites, shifts, masks, re, im, spin_inversion or 0)

    def state_info(self, alphas) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
        n = alphas.shape[0]
        betas = np.zeros(n, dtype=np.uint64)
        chars = np.zeros(n, dtype=np.complex128)
        norms = np.zeros(n, dtype=np.float64)
        lib().oracle_state_info(C.byref(self.c), n, _p(alphas, u64_p), _p(betas, u64_p),
                                C.cast(chars.ctypes.data, f64_p), _p(norms, f64_p))
        return betas, chars, norms

    def is_representative(self, alphas) -> Tuple[np.ndarray, np.ndarray]:
        alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
        n = alphas.shape[0]
        flags = np.zeros(n, dtype=np.uint8)
        norms = np.zeros(n, dtype=np.float64)
        lib().oracle_is_representative(C.byref(self.c), n, _p(alphas, u64_p), _p(flags, u8_p), _p(norms, f64_p))
        return flags, norms


class Basis:
    """Mirror of ls_hs_basis for the oracle (particle_type: 0 spin, 1 spinful, 2 spinless)."""

    def __init__(self, number_sites: int, particle_type: int = 0, number_particles: Optional[int] = None,
                 number_up: Optional[int] = None, spin_inversion: Optional[int] = None,
                 group: Optional[Group] = None):
        self.group = group
        b = oracle_basis()
        b.number_sites = number_sites
        b.number_particles = -1 if number_particles is None else number_particles
        b.number_up = -1 if number_up is None else number_up
        b.particle_type = particle_type
        b.spin_inversion = int(spin_inversion or 0)
        b.has_permutation_symmetries = int(group is not None and group.c.number_masks > 0 and particle_type == 0)
        if group is not None:
            b.group = group.c
        self.c = b
        self.number_bits = (2 if particle_type == 1 else 1) * number_sites

    @classmethod
    def from_product(cls, basis) -> "Basis":
        """Describe a lattice_symmetries_b200 basis object to the oracle."""
        group = None
        if basis._particle_type == 0:
            group = Group.from_symmetries(basis._symmetries, basis._number_sites, basis._spin_inversion)
        return cls(basis._number_sites, basis._particle_type, basis._number_particles, basis._number_up,
                   basis._spin_inversion, group)

    def min_state(self) -> int:
        return int(lib().oracle_min_state_estimate(C.byref(self.c)))

    def max_state(self) -> int:
        return int(lib().oracle_max_state_estimate(C.byref(self.c)))

    def enumerate(self) -> np.ndarray:
        count = C.c_uint64()
        p = lib().oracle_enumerate_states(C.byref(self.c), C.byref(count))
        return self._take(p, count.value)

    def enumerate_range(self, lower: int, upper: int) -> np.ndarray:
        """Representatives among the states of [lower, upper] (both of the basis'
        Hamming weight), threaded over chunks like StatesEnumeration.chpl:392-458."""
        count = C.c_uint64()
        p = lib().oracle_enumerate_range(C.byref(self.c), lower, upper, C.byref(count))
        return self._take(p, count.value)

    @staticmethod
    def _take(p, n: int) -> np.ndarray:
        if n == 0:
            out = np.zeros(0, dtype=np.uint64)
        else:
            out = np.frombuffer((C.c_uint64 * n).from_address(p), dtype=np.uint64).copy()
        if p:
            lib().oracle_free(p)
        return out


class Index:
    def __init__(self, reps: np.ndarray, number_bits: int, prefix_bits: int = 22):
        self.reps = np.ascontiguousarray(reps, dtype=np.uint64)
        self.prefix_bits = min(int(prefix_bits), int(number_bits))
        self.handle = lib().oracle_index_create(_p(self.reps, u64_p), self.reps.shape[0], number_bits, prefix_bits)

    def __call__(self, spins) -> np.ndarray:
        spins = np.ascontiguousarray(spins, dtype=np.uint64)
        out = np.zeros(spins.shape[0], dtype=np.int64)
        lib().oracle_state_index(self.handle, spins.shape[0], _p(spins, u64_p), _p(out, i64_p))
        return out

    def __del__(self):
        if getattr(self, "handle", None):
            lib().oracle_index_destroy(self.handle)
            self.handle = None


class Terms:
    def __init__(self, terms: Sequence):
        """``terms``: objects with v (complex), m, l, r, x, s (ints)."""
        self.n = len(terms)
        self.v = np.array([[t.v.real, t.v.imag] for t in terms], dtype=np.float64).reshape(self.n, 2)
        self.cols = {k: np.array([getattr(t, k) for t in terms], dtype=np.uint64) for k in "mlrxs"}
        c = oracle_terms()
        c.number_terms = self.n
        c.v = self.v.ctypes.data if self.n else None
        for k, a in self.cols.items():
            setattr(c, k, a.ctypes.data if self.n else None)
        self.c = c

    def ptr(self):
        return C.byref(self.c) if self.n > 0 else None


def apply_diag(terms: Terms, alphas, xs=None) -> np.ndarray:
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    ys = np.zeros(alphas.shape[0], dtype=np.float64)
    xs_p = None if xs is None else _p(np.ascontiguousarray(xs, dtype=np.float64), f64_p)
    lib().oracle_apply_diag(terms.ptr(), alphas.shape[0], _p(alphas, u64_p), _p(ys, f64_p), xs_p)
    return ys


def apply_off_diag(terms: Terms, alphas, xs=None):
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    cap = max(1, n * max(terms.n, 1))
    betas = np.zeros(cap, dtype=np.uint64)
    coeffs = np.zeros(cap, dtype=np.complex128)
    offsets = np.zeros(n + 1, dtype=np.int64)
    xs_arr = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
    lib().oracle_apply_off_diag(terms.ptr(), n, _p(alphas, u64_p), _p(betas, u64_p),
                                C.cast(coeffs.ctypes.data, f64_p), _p(offsets, i64_p),
                                None if xs_arr is None else _p(xs_arr, f64_p))
    total = int(offsets[-1])
    return betas[:total], coeffs[:total], offsets


def matvec(basis: Basis, off: Terms, diag: Terms, index: Index, x: np.ndarray, row_begin: int = 0,
           row_end: Optional[int] = None, sampling_prefix: bool = False, block_stride: int = 64,
           reference_kernels: bool = False) -> Tuple[np.ndarray, int]:
    """Push-form y = H x as the reference assembles it; returns (y, number of
    off-diagonal matrix elements emitted from the columns [row_begin, row_end)).

    ``block_stride`` > 64 (timing only): one block of 64 columns every ``block_stride`` columns -- a uniform sample.
    ``reference_kernels`` (timing only): apply_off_diag and state_index are the reference's own compiled
    kernels/reference.c and kernels/indexing.c (oracle/_ref/libref.so) instead of the restatement."""
    reps = index.reps
    dim = reps.shape[0]
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(dim, dtype=np.float64)
    if row_end is None:
        row_end = dim
    hooks = None
    if reference_kernels:
        R = ref()
        op, keep = _ref_op(off, diag, basis.number_bits)
        arr = _ext_array(reps.ctypes.data, dim, None)
        data = R.ls_hs_create_state_index_binary_search_kernel_data(C.byref(arr), basis.number_bits, index.prefix_bits)
        hooks = (op, keep, arr, data)
        lib().oracle_set_reference_hooks(
            C.cast(R.ls_internal_operator_apply_off_diag_x1, C.c_void_p), C.cast(C.pointer(op), C.c_void_p),
            C.cast(R.ls_hs_state_index_binary_search_kernel, C.c_void_p), data)
    try:
        n = lib().oracle_matvec_strided(C.byref(basis.c), off.ptr(), diag.ptr(), index.handle, _p(reps, u64_p), dim,
                                        row_begin, row_end, int(block_stride), _p(x, f64_p), _p(y, f64_p),
                                        3 if sampling_prefix else 1)
    finally:
        if hooks is not None:
            lib().oracle_set_reference_hooks(None, None, None, None)
            ref().ls_hs_destroy_state_index_binary_search_kernel_data(hooks[3])
    if n < 0:
        raise RuntimeError("oracle_matvec: invalid index (operator does not respect the basis symmetries)")
    return y, int(n)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


# ---- the compiled reference sources (kernels/indexing.c, kernels/reference.c) -----
class _ext_array(C.Structure):
    _fields_ = [("elts", C.c_void_p), ("num_elts", C.c_uint64), ("freer", C.c_void_p)]


class _ref_terms(C.Structure):  # kernels/lattice_symmetries_types.h:140-151
    _fields_ = [("number_terms", C.c_int), ("number_bits", C.c_int), ("v", C.c_void_p), ("m", C.c_void_p),
                ("l", C.c_void_p), ("r", C.c_void_p), ("x", C.c_void_p), ("s", C.c_void_p)]


class _ref_operator(C.Structure):  # :153-161
    _fields_ = [("refcount", C.c_int), ("basis", C.c_void_p), ("off_diag_terms", C.POINTER(_ref_terms)),
                ("diag_terms", C.POINTER(_ref_terms)), ("haskell_payload", C.c_void_p)]


_ref = None


def ref_available() -> bool:
    build()
    return LIBREF.exists()


def ref():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(str(LIBREF))
        R.ls_hs_create_state_index_binary_search_kernel_data.argtypes = [C.POINTER(_ext_array), C.c_int, C.c_int]
        R.ls_hs_create_state_index_binary_search_kernel_data.restype = C.c_void_p
        R.ls_hs_destroy_state_index_binary_search_kernel_data.argtypes = [C.c_void_p]
        R.ls_hs_state_index_binary_search_kernel.argtypes = [C.c_ssize_t, u64_p, C.c_ssize_t, i64_p, C.c_ssize_t,
                                                             C.c_void_p]
        R.ls_internal_operator_apply_diag_x1.argtypes = [C.POINTER(_ref_operator), C.c_ssize_t, u64_p, f64_p, f64_p]
        R.ls_internal_operator_apply_off_diag_x1.argtypes = [C.POINTER(_ref_operator), C.c_ssize_t, u64_p, u64_p,
                                                             C.c_void_p, i64_p, f64_p]
        _ref = R
    return _ref


def ref_state_index(reps: np.ndarray, number_bits: int, prefix_bits: int, spins: np.ndarray) -> np.ndarray:
    """kernels/indexing.c through its own entry points."""
    reps = np.ascontiguousarray(reps, dtype=np.uint64)
    spins = np.ascontiguousarray(spins, dtype=np.uint64)
    arr = _ext_array(reps.ctypes.data, reps.shape[0], None)
    R = ref()
    data = R.ls_hs_create_state_index_binary_search_kernel_data(C.byref(arr), number_bits, prefix_bits)
    out = np.zeros(spins.shape[0], dtype=np.int64)
    R.ls_hs_state_index_binary_search_kernel(spins.shape[0], _p(spins, u64_p), 1, _p(out, i64_p), 1, data)
    R.ls_hs_destroy_state_index_binary_search_kernel_data(data)
    return out


def _ref_op(off: Terms, diag: Terms, number_bits: int):
    keep = []

    def mk(t: Terms):
        if t.n == 0:
            return None
        s = _ref_terms(t.n, number_bits, t.v.ctypes.data, *[t.cols[k].ctypes.data for k in "mlrxs"])
        keep.append(s)
        return C.pointer(s)

    op = _ref_operator()
    op.refcount = 1
    o, d = mk(off), mk(diag)
    if o is not None:
        op.off_diag_terms = o
    if d is not None:
        op.diag_terms = d
    keep.append(op)
    return op, keep


def ref_apply_diag(off: Terms, diag: Terms, number_bits: int, alphas, xs=None) -> np.ndarray:
    """kernels/reference.c:67-95 through its own entry point."""
    op, keep = _ref_op(off, diag, number_bits)
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    ys = np.zeros(alphas.shape[0], dtype=np.float64)
    xs_arr = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
    ref().ls_internal_operator_apply_diag_x1(C.byref(op), alphas.shape[0], _p(alphas, u64_p), _p(ys, f64_p),
                                             None if xs_arr is None else _p(xs_arr, f64_p))
    return ys


def ref_apply_off_diag(off: Terms, diag: Terms, number_bits: int, alphas, xs=None):
    """kernels/reference.c:97-134 through its own entry point."""
    op, keep = _ref_op(off, diag, number_bits)
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    cap = max(1, n * max(off.n, 1))
    betas = np.zeros(cap, dtype=np.uint64)
    coeffs = np.zeros(cap, dtype=np.complex128)
    offsets = np.zeros(n + 1, dtype=np.int64)
    xs_arr = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
    ref().ls_internal_operator_apply_off_diag_x1(C.byref(op), n, _p(alphas, u64_p), _p(betas, u64_p),
                                                 coeffs.ctypes.data, _p(offsets, i64_p),
                                                 None if xs_arr is None else _p(xs_arr, f64_p))
    total = int(offsets[-1])
    return betas[:total], coeffs[:total], offsets
