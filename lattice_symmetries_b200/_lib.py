"""ctypes binding of liblattice_symmetries_b200.so (the C ABI declared in
include/lattice_symmetries_b200.h).

This plays the role of the reference's cffi module
(python/lattice_symmetries/build_extension.py:83-107, which links
``-llattice_symmetries_chapel -llattice_symmetries_haskell``): struct layouts
are those of kernels/lattice_symmetries_types.h, and ``ls_chpl_init`` registers
the kernel vtable exactly as python/lattice_symmetries/__init__.py:50-55 expects.

There is no CPU fallback: a missing library is an ImportError, and every
compute entry point fails through ``ls_hs_error`` (-> RuntimeError, like
build_extension.py:70-81) when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

__all__ = ["lib", "LIB_PATH"]

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LS_B200_LIBRARY", _HERE / "liblattice_symmetries_b200.so"))

u64_p = C.POINTER(C.c_uint64)
f64_p = C.POINTER(C.c_double)
i64_p = C.POINTER(C.c_int64)
u8_p = C.POINTER(C.c_uint8)


class chpl_external_array(C.Structure):
    _fields_ = [("elts", C.c_void_p), ("num_elts", C.c_uint64), ("freer", C.c_void_p)]


class ls_hs_scalar(C.Structure):
    _fields_ = [("_real", C.c_double), ("_imag", C.c_double)]


class ls_hs_basis_kernels(C.Structure):
    _fields_ = [
        ("state_info_kernel", C.c_void_p),
        ("state_info_data", C.c_void_p),
        ("is_representative_kernel", C.c_void_p),
        ("is_representative_data", C.c_void_p),
        ("state_index_kernel", C.c_void_p),
        ("state_index_data", C.c_void_p),
    ]


class ls_hs_permutation_group(C.Structure):
    _fields_ = [
        ("refcount", C.c_int),
        ("number_bits", C.c_int),
        ("number_shifts", C.c_int),
        ("number_masks", C.c_int),
        ("masks", C.c_void_p),
        ("shifts", C.c_void_p),
        ("eigvals_re", C.c_void_p),
        ("eigvals_im", C.c_void_p),
        ("haskell_payload", C.c_void_p),
    ]


class ls_hs_basis(C.Structure):
    _fields_ = [
        ("refcount", C.c_int),
        ("number_sites", C.c_int),
        ("number_particles", C.c_int),
        ("number_up", C.c_int),
        ("particle_type", C.c_int),
        ("spin_inversion", C.c_int),
        ("state_index_is_identity", C.c_bool),
        ("requires_projection", C.c_bool),
        ("kernels", C.POINTER(ls_hs_basis_kernels)),
        ("representatives", chpl_external_array),
        ("haskell_payload", C.c_void_p),
    ]


class ls_hs_nonbranching_terms(C.Structure):
    _fields_ = [
        ("number_terms", C.c_int),
        ("number_bits", C.c_int),
        ("v", C.c_void_p),
        ("m", C.c_void_p),
        ("l", C.c_void_p),
        ("r", C.c_void_p),
        ("x", C.c_void_p),
        ("s", C.c_void_p),
    ]


class ls_hs_operator(C.Structure):
    _fields_ = [
        ("refcount", C.c_int),
        ("basis", C.POINTER(ls_hs_basis)),
        ("off_diag_terms", C.POINTER(ls_hs_nonbranching_terms)),
        ("diag_terms", C.POINTER(ls_hs_nonbranching_terms)),
        ("haskell_payload", C.c_void_p),
    ]


ENUMERATE_STATES = C.CFUNCTYPE(None, C.POINTER(ls_hs_basis), C.c_uint64, C.c_uint64, C.POINTER(chpl_external_array))
APPLY_OFF_DIAG = C.CFUNCTYPE(
    None, C.POINTER(ls_hs_operator), C.c_int64, u64_p, C.POINTER(chpl_external_array),
    C.POINTER(chpl_external_array), C.POINTER(chpl_external_array), C.c_int64)
APPLY_DIAG = C.CFUNCTYPE(None, C.POINTER(ls_hs_operator), C.c_int64, u64_p, C.POINTER(chpl_external_array), C.c_int64)
MATVEC = C.CFUNCTYPE(None, C.POINTER(ls_hs_operator), C.c_int, f64_p, f64_p)


class ls_chpl_kernels(C.Structure):
    _fields_ = [
        ("enumerate_states", ENUMERATE_STATES),
        ("operator_apply_off_diag", APPLY_OFF_DIAG),
        ("operator_apply_diag", APPLY_DIAG),
        ("matrix_vector_product", MATVEC),
    ]


ERROR_HANDLER = C.CFUNCTYPE(None, C.c_char_p)

# name -> (restype, argtypes): every symbol include/lattice_symmetries_b200.h declares
PROTOTYPES = {
    "ls_hs_set_exception_handler": (None, [ERROR_HANDLER]),
    "ls_hs_error": (None, [C.c_char_p]),
    "ls_hs_fatal_error": (None, [C.c_char_p, C.c_int, C.c_char_p]),
    "ls_hs_internal_destroy_external_array": (None, [C.POINTER(chpl_external_array)]),
    "ls_hs_internal_read_refcount": (C.c_int, [C.POINTER(C.c_int)]),
    "ls_hs_internal_write_refcount": (None, [C.POINTER(C.c_int), C.c_int]),
    "ls_hs_internal_inc_refcount": (C.c_int, [C.POINTER(C.c_int)]),
    "ls_hs_internal_dec_refcount": (C.c_int, [C.POINTER(C.c_int)]),
    "ls_internal_create_halide_kernel_data": (C.c_void_p, [C.POINTER(ls_hs_permutation_group), C.c_int]),
    "ls_internal_destroy_halide_kernel_data": (None, [C.c_void_p]),
    "ls_hs_is_representative_halide_kernel": (None, [C.c_ssize_t, u64_p, C.c_ssize_t, u8_p, f64_p, C.c_void_p]),
    "ls_hs_state_info_halide_kernel": (
        None, [C.c_ssize_t, u64_p, C.c_ssize_t, u64_p, C.c_ssize_t, C.c_void_p, f64_p, C.c_void_p]),
    "ls_hs_create_state_index_binary_search_kernel_data": (
        C.c_void_p, [C.POINTER(chpl_external_array), C.c_int, C.c_int]),
    "ls_hs_destroy_state_index_binary_search_kernel_data": (None, [C.c_void_p]),
    "ls_hs_state_index_binary_search_kernel": (None, [C.c_ssize_t, u64_p, C.c_ssize_t, i64_p, C.c_ssize_t, C.c_void_p]),
    "ls_hs_state_index": (None, [C.POINTER(ls_hs_basis), C.c_ssize_t, u64_p, C.c_ssize_t, i64_p, C.c_ssize_t]),
    "ls_hs_is_representative": (None, [C.POINTER(ls_hs_basis), C.c_ssize_t, u64_p, C.c_ssize_t, u8_p, f64_p]),
    "ls_hs_state_info": (
        None, [C.POINTER(ls_hs_basis), C.c_ssize_t, u64_p, C.c_ssize_t, u64_p, C.c_ssize_t, C.c_void_p, f64_p]),
    "ls_hs_build_representatives": (None, [C.POINTER(ls_hs_basis), C.c_uint64, C.c_uint64]),
    "ls_hs_unchecked_set_representatives": (None, [C.POINTER(ls_hs_basis), C.POINTER(chpl_external_array), C.c_int]),
    "ls_internal_operator_apply_diag_x1": (None, [C.POINTER(ls_hs_operator), C.c_ssize_t, u64_p, f64_p, f64_p]),
    "ls_internal_operator_apply_off_diag_x1": (
        None, [C.POINTER(ls_hs_operator), C.c_ssize_t, u64_p, u64_p, C.c_void_p, i64_p, f64_p]),
    "ls_hs_internal_get_chpl_kernels": (C.POINTER(ls_chpl_kernels), []),
    "ls_hs_internal_set_chpl_kernels": (None, [C.POINTER(ls_chpl_kernels)]),
    "ls_chpl_init": (None, []),
    "ls_chpl_finalize": (None, []),
    "ls_chpl_init_kernels": (None, []),
    "ls_chpl_enumerate_representatives": (
        None, [C.POINTER(ls_hs_basis), C.c_uint64, C.c_uint64, C.POINTER(chpl_external_array)]),
    "ls_chpl_operator_apply_diag": (
        None, [C.POINTER(ls_hs_operator), C.c_int64, u64_p, C.POINTER(chpl_external_array), C.c_int64]),
    "ls_chpl_operator_apply_off_diag": (
        None, [C.POINTER(ls_hs_operator), C.c_int64, u64_p, C.POINTER(chpl_external_array),
               C.POINTER(chpl_external_array), C.POINTER(chpl_external_array), C.c_int64]),
    "ls_chpl_matrix_vector_product": (None, [C.POINTER(ls_hs_operator), C.c_int, f64_p, f64_p]),
    # extensions (device-resident path)
    "ls_b200_kernel_launch_count": (C.c_uint64, []),
    "ls_b200_last_kernel_ms": (C.c_double, [C.c_char_p]),
    "ls_b200_stream": (C.c_void_p, []),
    "ls_b200_set_stream": (C.c_int, [C.c_void_p]),
    "ls_b200_device_count": (C.c_int, []),
    "ls_b200_measure_lop3_peak": (C.c_double, []),
    "ls_b200_basis_device_view": (
        C.c_int, [C.POINTER(ls_hs_basis), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ls_b200_index_info": (C.c_int, [C.POINTER(ls_hs_basis), C.POINTER(C.c_int64)]),
    "ls_b200_matvec_device": (C.c_int, [C.POINTER(ls_hs_operator), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ls_b200_matvec_device_phase": (
        C.c_int, [C.POINTER(ls_hs_operator), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "ls_b200_matvec_block_device": (
        C.c_int, [C.POINTER(ls_hs_operator), C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "ls_b200_matvec_device_c128": (C.c_int, [C.POINTER(ls_hs_operator), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ls_b200_matvec_sync": (C.c_int, []),
    "ls_b200_count_matrix_elements": (C.c_int64, [C.POINTER(ls_hs_operator), C.c_int64, C.c_int64]),
    "ls_b200_build_shard": (
        C.c_int, [C.POINTER(ls_hs_basis), C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                  C.POINTER(C.c_uint64)]),
    "ls_b200_build_blocks": (
        C.c_int, [C.POINTER(ls_hs_basis), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p),
                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ls_b200_number_candidates": (C.c_uint64, [C.POINTER(ls_hs_basis)]),
    "ls_b200_set_representatives_device": (
        C.c_int, [C.POINTER(ls_hs_basis), C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "ls_b200_device_malloc": (C.c_void_p, [C.c_size_t]),
    "ls_b200_device_free": (None, [C.c_void_p]),
    "ls_b200_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ls_b200_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ls_b200_host_malloc": (C.c_void_p, [C.c_size_t]),
    "ls_b200_host_free": (None, [C.c_void_p]),
    "ls_b200_operator_release": (None, [C.c_void_p]),
    "ls_b200_operator_apply_off_diag_projected": (
        C.c_int64, [C.POINTER(ls_hs_operator), C.c_int64, u64_p, u64_p, C.c_void_p, i64_p, i64_p]),
    # several GPUs of one node (dist.cu)
    "ls_b200_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "ls_b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "ls_b200_comm_finalize": (None, []),
    "ls_b200_comm_size": (C.c_int, []),
    "ls_b200_comm_rank": (C.c_int, []),
    "ls_b200_comm_allreduce_f64": (C.c_int, [C.c_void_p, C.c_int]),
    "ls_b200_dist_build": (C.c_int, [C.POINTER(ls_hs_basis), C.c_void_p, C.c_int]),
    "ls_b200_dist_matvec": (C.c_int, [C.POINTER(ls_hs_operator), C.c_void_p, C.c_void_p, C.c_int]),
    "ls_b200_dist_matvec_c128": (C.c_int, [C.POINTER(ls_hs_operator), C.c_void_p, C.c_void_p, C.c_int]),
    "ls_b200_dist_rebalance": (C.c_int, [C.POINTER(ls_hs_basis)]),
    "ls_b200_emu_rebalance": (C.c_int, [C.c_void_p, C.c_int, f64_p]),
    "ls_b200_dist_info": (C.c_int, [C.POINTER(ls_hs_basis), C.POINTER(C.c_int64)]),
    "ls_b200_dist_bounds": (C.c_int, [C.POINTER(ls_hs_basis), C.POINTER(C.c_int64), C.c_int]),
    "ls_b200_emu_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "ls_b200_emu_matvec": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "ls_b200_plan_blocks": (C.c_int64, [C.c_uint64, C.c_int, u64_p, u64_p, C.c_int64]),
    "ls_b200_plan_redistribution": (
        C.c_int64, [C.c_int, C.c_int, C.c_int64, i64_p, C.POINTER(C.c_int32), i64_p, i64_p, i64_p, i64_p, i64_p, i64_p,
                    C.c_int64]),
    "ls_b200_plan_balanced_bounds": (C.c_int, [C.c_int64, i64_p, f64_p, C.c_int, i64_p]),
    "ls_b200_plan_rebalance_bounds": (C.c_int, [C.c_int64, i64_p, f64_p, C.c_int, C.c_int64, i64_p]),
    "ls_b200_plan_cyclic_share": (C.c_int, [C.c_uint64, C.c_int, C.c_int, u64_p]),
}


class DeviceArray:
    """A device allocation owned through the C ABI (ls_b200_device_malloc);
    numpy in / numpy out.  Host-side plumbing for tests, bench and the Lanczos
    driver -- no torch types cross the boundary."""

    def __init__(self, count: int, dtype):
        import numpy as np
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        self.ptr = lib.ls_b200_device_malloc(max(self.nbytes, 8))
        check_error()
        if not self.ptr:
            raise MemoryError(f"ls_b200_device_malloc({self.nbytes}) failed")

    @property
    def nbytes(self) -> int:
        return self.count * self.dtype.itemsize

    @classmethod
    def from_numpy(cls, a) -> "DeviceArray":
        import numpy as np
        a = np.ascontiguousarray(a)
        d = cls(a.size, a.dtype)
        d.upload(a)
        return d

    def upload(self, a) -> None:
        import numpy as np
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size == self.count
        lib.ls_b200_copy_to_device(self.ptr, a.ctypes.data, self.nbytes)
        check_error()

    def numpy(self):
        import numpy as np
        out = np.empty(self.count, dtype=self.dtype)
        lib.ls_b200_copy_to_host(out.ctypes.data, self.ptr, self.nbytes)
        check_error()
        return out

    def free(self) -> None:
        if self.ptr:
            lib.ls_b200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_to_numpy(ptr: int, count: int, dtype):
    """Copy ``count`` elements from a raw device pointer."""
    import numpy as np
    out = np.empty(int(count), dtype=np.dtype(dtype))
    if count:
        lib.ls_b200_copy_to_host(out.ctypes.data, ptr, out.nbytes)
        check_error()
    return out


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C lattice_symmetries_b200/csrc`); this package has no CPU fallback")
    handle = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return handle


lib = _load()

# -- error handling: build_extension.py:70-81 ------------------------------------
_pending_error: list = []


@ERROR_HANDLER
def _python_error_handler(message):
    _pending_error.append(message.decode("utf-8", "replace") if message else "unknown error")


lib.ls_hs_set_exception_handler(_python_error_handler)


def check_error() -> None:
    """Raise the RuntimeError recorded by the library's error handler, if any."""
    if _pending_error:
        msg = "; ".join(_pending_error)
        _pending_error.clear()
        raise RuntimeError(msg)


_initialised = False


def ensure_init() -> None:
    """python/lattice_symmetries/__init__.py:50-55: ls_chpl_init registers the
    kernel vtable (and here brings up the CUDA device)."""
    global _initialised
    if not _initialised:
        lib.ls_chpl_init()
        check_error()
        _initialised = True


def symbol_address(name: str) -> int:
    return C.cast(getattr(lib, name), C.c_void_p).value
