"""Hilbert-space bases: the host-side mirror of the reference's ``Basis`` API
(python/lattice_symmetries/__init__.py:148-399).

In the reference the ``ls_hs_basis`` struct is produced by the Haskell host
(haskell/src/LatticeSymmetries/Basis.hs:788-810, kernels wired at :927-997).
GHC is not available here, so this module fills the very same struct -- same
field values, same kernel symbols -- and hands it to the C ABI.  All compute
(``build``, ``index``, ``state_info``, ``is_representative``) runs in the CUDA
library; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib
from ._lib import lib
from .symmetry import Symmetries

__all__ = ["Basis", "SpinBasis", "SpinlessFermionBasis", "SpinfulFermionBasis"]

LS_HS_SPIN, LS_HS_SPINFUL_FERMION, LS_HS_SPINLESS_FERMION = 0, 1, 2


def _low_ones(h: int) -> int:
    return (1 << h) - 1


def _destroy_basis(keep: dict) -> None:
    """Basis.hs:999-1030 destroyCbasis_kernels + :815-822."""
    basis = keep["basis"]
    kernels = keep["kernels"]
    if kernels.state_index_data:
        lib.ls_hs_destroy_state_index_binary_search_kernel_data(kernels.state_index_data)
        kernels.state_index_data = None
    if kernels.state_info_data:
        lib.ls_internal_destroy_halide_kernel_data(kernels.state_info_data)
        kernels.state_info_data = None
        kernels.is_representative_data = None
    if basis.representatives.freer:
        lib.ls_hs_internal_destroy_external_array(C.byref(basis.representatives))
        basis.representatives.elts = None
        basis.representatives.num_elts = 0
        basis.representatives.freer = None


class Basis:
    """Common behaviour; construct one of the subclasses."""

    def __init__(self, *, number_sites: int, particle_type: int, number_particles: Optional[int],
                 number_up: Optional[int], spin_inversion: Optional[int], symmetries: Optional[Symmetries]):
        _lib.ensure_init()
        self._symmetries = symmetries if symmetries is not None else Symmetries([])
        self._number_sites = int(number_sites)
        self._particle_type = particle_type
        self._number_particles = number_particles
        self._number_up = number_up
        self._spin_inversion = spin_inversion
        if self.number_bits > 64:
            raise ValueError("bases with more than 64 bits are not supported")  # ForeignLibrary.hs:163-168
        if particle_type == LS_HS_SPIN and not self._symmetries.is_empty \
                and self._symmetries.number_bits != self._number_sites:
            raise ValueError("symmetries act on a different number of sites")

        kernels = _lib.ls_hs_basis_kernels()
        keep = {"kernels": kernels}
        if particle_type == LS_HS_SPIN:
            # Basis.hs:927-948: every spin basis gets kernel data, even for the empty group
            nbits, shifts, masks, re, im = self._symmetries.tables()
            group = _lib.ls_hs_permutation_group()
            group.refcount = 1
            group.number_bits = nbits if not self._symmetries.is_empty else self._number_sites
            group.number_shifts = len(shifts)
            group.number_masks = len(re)
            shifts = np.ascontiguousarray(shifts, dtype=np.uint64)
            group.masks = masks.ctypes.data if masks.size else None
            group.shifts = shifts.ctypes.data if shifts.size else None
            group.eigvals_re = re.ctypes.data if re.size else None
            group.eigvals_im = im.ctypes.data if im.size else None
            data = lib.ls_internal_create_halide_kernel_data(C.byref(group), int(spin_inversion or 0))
            _lib.check_error()
            kernels.state_info_kernel = _lib.symbol_address("ls_hs_state_info_halide_kernel")
            kernels.state_info_data = data
            kernels.is_representative_kernel = _lib.symbol_address("ls_hs_is_representative_halide_kernel")
            kernels.is_representative_data = data
        basis = _lib.ls_hs_basis()
        basis.refcount = 1
        basis.number_sites = self._number_sites
        basis.number_particles = -1 if number_particles is None else int(number_particles)
        basis.number_up = -1 if number_up is None else int(number_up)
        basis.particle_type = particle_type
        basis.spin_inversion = int(spin_inversion or 0)
        basis.state_index_is_identity = self._state_index_is_identity()
        basis.requires_projection = self.requires_projection
        basis.kernels = C.pointer(kernels)
        keep["basis"] = basis
        self._payload = basis
        self._keep = keep
        self._finalizer = weakref.finalize(self, _destroy_basis, keep)

    # -- Basis.hs:701-774 predicates ----------------------------------------
    def _state_index_is_identity(self) -> bool:
        if self._particle_type == LS_HS_SPIN:
            return self._number_up is None and self._spin_inversion is None and self._symmetries.is_empty
        return self._number_particles is None

    @property
    def number_sites(self) -> int:
        return self._number_sites

    @property
    def number_bits(self) -> int:
        return (2 if self._particle_type == LS_HS_SPINFUL_FERMION else 1) * self._number_sites

    @property
    def number_words(self) -> int:
        return (self.number_bits + 63) // 64

    @property
    def has_spin_inversion_symmetry(self) -> bool:
        return self._particle_type == LS_HS_SPIN and self._spin_inversion is not None

    @property
    def has_permutation_symmetries(self) -> bool:
        return self._particle_type == LS_HS_SPIN and not self._symmetries.is_empty

    @property
    def requires_projection(self) -> bool:
        return self.has_permutation_symmetries or self.has_spin_inversion_symmetry

    @property
    def has_fixed_hamming_weight(self) -> bool:
        if self._particle_type == LS_HS_SPIN:
            return self._number_up is not None
        return self._number_particles is not None

    @property
    def symmetries(self) -> Symmetries:
        return self._symmetries

    @property
    def min_state_estimate(self) -> int:
        n = self._number_sites
        if self._particle_type == LS_HS_SPINFUL_FERMION and self._number_up is not None:
            up, down = self._number_up, self._number_particles - self._number_up
            return (_low_ones(down) << n) | _low_ones(up)
        if self.has_fixed_hamming_weight:
            h = self._number_up if self._particle_type == LS_HS_SPIN else self._number_particles
            return _low_ones(h)
        return 0

    @property
    def max_state_estimate(self) -> int:
        n = self._number_sites
        if self._particle_type == LS_HS_SPINFUL_FERMION and self._number_up is not None:
            up, down = self._number_up, self._number_particles - self._number_up
            return ((_low_ones(down) << (n - down)) << n) | (_low_ones(up) << (n - up))
        bits = self.number_bits
        if self.has_fixed_hamming_weight:
            h = self._number_up if self._particle_type == LS_HS_SPIN else self._number_particles
            if self.has_spin_inversion_symmetry:
                return 0 if h == 0 else _low_ones(h) << (bits - h - 1)
            return _low_ones(h) << (bits - h)
        return _low_ones(bits)

    # -- text forms (python/lattice_symmetries/__init__.py:248-254, 320-329) --------
    def state_to_string(self, state: int) -> str:
        """Pretty-print a basis state."""
        from .config import state_to_string
        return state_to_string(state, self.number_bits, self._particle_type == LS_HS_SPINFUL_FERMION)

    def to_json(self) -> str:
        import json
        from .config import basis_header
        particle = {LS_HS_SPIN: "spin-1/2", LS_HS_SPINFUL_FERMION: "spinful-fermion",
                    LS_HS_SPINLESS_FERMION: "spinless-fermion"}[self._particle_type]
        occupation = self._number_particles
        if self._particle_type == LS_HS_SPINFUL_FERMION and self._number_up is not None:
            occupation = (self._number_up, self._number_particles - self._number_up)
        return json.dumps(basis_header(particle, self._number_sites, self._number_up, self._spin_inversion,
                                       self._symmetries, occupation))

    @staticmethod
    def from_json(json_string: str) -> "Basis":
        import json
        from .config import parse_config
        if not isinstance(json_string, str):
            raise TypeError(f"expected a str, got {type(json_string).__name__}")
        return parse_config({"basis": json.loads(json_string)}).model.basis()

    # -- build / queries ---------------------------------------------------------
    @property
    def is_built(self) -> bool:
        return bool(self._payload.kernels.contents.state_index_kernel)

    def check_is_built(self) -> None:
        if not self.is_built:
            raise ValueError(
                "basis states have not been built yet; "
                "if you wish to do so, use the basis.build() function")

    def build(self) -> None:
        """Generate the list of representatives on the GPU
        (reference: Basis.build -> ls_hs_basis_build -> ls_hs_build_representatives)."""
        if not self.is_built:
            lib.ls_hs_build_representatives(C.byref(self._payload), self.min_state_estimate, self.max_state_estimate)
            _lib.check_error()
        assert self.is_built

    def unchecked_set_representatives(self, states) -> None:
        """kernels/reference.c:196-211 (borrows ``states``)."""
        states = np.ascontiguousarray(states, dtype=np.uint64)
        self._keep["borrowed_states"] = states
        arr = _lib.chpl_external_array(states.ctypes.data, states.size, None)
        lib.ls_hs_unchecked_set_representatives(C.byref(self._payload), C.byref(arr), 22)
        _lib.check_error()

    @property
    def number_states(self) -> int:
        self.check_is_built()
        return int(self._payload.representatives.num_elts)

    @property
    def states(self) -> np.ndarray:
        n = self.number_states
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        buf = (C.c_uint64 * n).from_address(self._payload.representatives.elts)
        arr = np.frombuffer(buf, dtype=np.uint64)
        arr.flags.writeable = False
        self._keep.setdefault("views", []).append(buf)
        return arr

    def state_info(self, x):
        """Representative, character and norm of basis states
        (python/lattice_symmetries/__init__.py:256-297)."""
        assert self.number_bits <= 64
        is_scalar = isinstance(x, (int, np.integer))
        x = np.array([x], dtype=np.uint64) if is_scalar else np.ascontiguousarray(x, dtype=np.uint64)
        count = x.shape[0]
        if self.has_permutation_symmetries:
            betas = np.zeros_like(x)
            characters = np.zeros(count, dtype=np.complex128)
            norms = np.zeros(count, dtype=np.float64)
            lib.ls_hs_state_info(
                C.byref(self._payload), count, x.ctypes.data_as(_lib.u64_p), 1, betas.ctypes.data_as(_lib.u64_p), 1,
                characters.ctypes.data, norms.ctypes.data_as(_lib.f64_p))
            _lib.check_error()
        elif self.has_spin_inversion_symmetry:
            mask = (1 << self.number_bits) - 1
            betas = np.bitwise_xor(x, np.uint64(mask))
            when = betas < x
            betas = np.where(when, betas, x)
            characters = np.where(when, float(self._spin_inversion), 1.0).astype(np.complex128)
            norms = np.ones(count, dtype=np.float64)
        else:
            betas = x
            characters = np.ones(count, dtype=np.complex128)
            norms = np.ones(count, dtype=np.float64)
        if is_scalar:
            return (int(betas[0]), complex(characters[0]), float(norms[0]))
        return (betas, characters, norms)

    def is_representative(self, x) -> Tuple[np.ndarray, np.ndarray]:
        """Flags and raw stabiliser sums (ls_hs_is_representative,
        kernels/reference.c:150-159); norms are meaningful where the flag is set."""
        x = np.ascontiguousarray(x, dtype=np.uint64)
        flags = np.zeros(x.shape[0], dtype=np.uint8)
        norms = np.zeros(x.shape[0], dtype=np.float64)
        lib.ls_hs_is_representative(
            C.byref(self._payload), x.shape[0], x.ctypes.data_as(_lib.u64_p), 1, flags.ctypes.data_as(_lib.u8_p),
            norms.ctypes.data_as(_lib.f64_p))
        _lib.check_error()
        return flags, norms

    def index(self, x) -> Union[int, np.ndarray]:
        """Index of basis states, -1 when absent
        (python/lattice_symmetries/__init__.py:299-318)."""
        self.check_is_built()
        x = np.asarray(x, dtype=np.uint64, order="C")
        is_scalar = x.ndim == 0
        if is_scalar:
            x = np.expand_dims(x, axis=0)
        x = np.ascontiguousarray(x)
        indices = np.zeros(x.shape[0], dtype=np.int64)
        lib.ls_hs_state_index(
            C.byref(self._payload), x.shape[0], x.ctypes.data_as(_lib.u64_p), 1, indices.ctypes.data_as(_lib.i64_p), 1)
        _lib.check_error()
        return int(indices[0]) if is_scalar else indices

    # -- device-resident extensions ------------------------------------------------
    def index_info(self) -> dict:
        """Shape of the device index: prefix bits, search trip count, second-level tables."""
        self.check_is_built()
        out = (C.c_int64 * 4)()
        if lib.ls_b200_index_info(C.byref(self._payload), out) != 0:
            raise RuntimeError("basis has no device index")
        return {"prefix_bits": int(out[0]), "steps": int(out[1]), "two_level": bool(out[2]), "states": int(out[3])}

    @property
    def number_candidates(self) -> int:
        """Size of the enumeration range scanned by ``build`` (combinadic index space)."""
        n = int(lib.ls_b200_number_candidates(C.byref(self._payload)))
        _lib.check_error()
        return n

    def device_view(self) -> Tuple[int, int, int]:
        """(device pointer to sorted representatives, device pointer to norms or 0, count)."""
        reps, norms, count = C.c_void_p(), C.c_void_p(), C.c_uint64()
        status = lib.ls_b200_basis_device_view(C.byref(self._payload), C.byref(reps), C.byref(norms), C.byref(count))
        _lib.check_error()
        if status != 0:
            raise RuntimeError("basis is not built")
        return (reps.value or 0, norms.value or 0, int(count.value))

    def build_shard(self, index_begin: int, index_end: int) -> Tuple[int, int, int]:
        """Scan candidates [index_begin, index_end) and return (reps_dev, norms_dev, count);
        the caller owns the device buffers (``lib.ls_b200_device_free``)."""
        reps, norms, count = C.c_void_p(), C.c_void_p(), C.c_uint64()
        status = lib.ls_b200_build_shard(
            C.byref(self._payload), int(index_begin), int(index_end), C.byref(reps), C.byref(norms), C.byref(count))
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_build_shard failed")
        return (reps.value or 0, norms.value or 0, int(count.value))

    def build_blocks(self, first_begin: int, block_size: int, stride: int, number_blocks: int):
        """Scan the candidate blocks [first_begin + k stride, + block_size), k < number_blocks, in one call;
        returns (reps_dev, norms_dev, counts per block); the caller owns the device buffers."""
        reps, norms = C.c_void_p(), C.c_void_p()
        counts = (C.c_uint64 * max(int(number_blocks), 1))()
        status = lib.ls_b200_build_blocks(
            C.byref(self._payload), int(first_begin), int(block_size), int(stride), int(number_blocks),
            C.byref(reps), C.byref(norms), counts)
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_build_blocks failed")
        return (reps.value or 0, norms.value or 0, [int(c) for c in counts[:int(number_blocks)]])

    def set_representatives_device(self, reps_dev: int, norms_dev: int, count: int, cache_bits: int = 22) -> None:
        status = lib.ls_b200_set_representatives_device(
            C.byref(self._payload), reps_dev, norms_dev or None, int(count), cache_bits)
        _lib.check_error()
        if status != 0:
            raise RuntimeError("ls_b200_set_representatives_device failed")


class SpinBasis(Basis):
    def __init__(self, number_spins: int, hamming_weight: Optional[int] = None,
                 spin_inversion: Optional[int] = None, symmetries: Optional[Symmetries] = None):
        """Hilbert space basis for ``number_spins`` spin-1/2 particles
        (python/lattice_symmetries/__init__.py:330-359)."""
        if spin_inversion not in (None, 1, -1):
            raise ValueError(f"invalid spin_inversion: {spin_inversion}; expected 1, -1 or None")
        if hamming_weight is not None and not (0 <= hamming_weight <= number_spins):
            raise ValueError(f"invalid hamming_weight: {hamming_weight}")
        if spin_inversion is not None and hamming_weight is not None and 2 * hamming_weight != number_spins:
            raise ValueError("spin inversion requires hamming_weight == number_spins / 2")  # Basis.hs:245-256
        super().__init__(number_sites=number_spins, particle_type=LS_HS_SPIN, number_particles=number_spins,
                         number_up=hamming_weight, spin_inversion=spin_inversion, symmetries=symmetries)

    @property
    def spin_inversion(self) -> Optional[int]:
        return self._spin_inversion


class SpinlessFermionBasis(Basis):
    def __init__(self, number_sites: int, number_particles: Optional[int] = None):
        """python/lattice_symmetries/__init__.py:362-380."""
        super().__init__(number_sites=number_sites, particle_type=LS_HS_SPINLESS_FERMION,
                         number_particles=number_particles, number_up=None, spin_inversion=None, symmetries=None)


class SpinfulFermionBasis(Basis):
    def __init__(self, number_sites: int, number_particles: Union[None, int, Tuple[int, int]] = None):
        """python/lattice_symmetries/__init__.py:383-402: ``number_particles`` is
        None, a total, or ``(N_up, N_down)``."""
        if isinstance(number_particles, (tuple, list)):
            up, down = int(number_particles[0]), int(number_particles[1])
            total, n_up = up + down, up
        elif number_particles is None:
            total, n_up = None, None
        else:
            total, n_up = int(number_particles), None
        super().__init__(number_sites=number_sites, particle_type=LS_HS_SPINFUL_FERMION,
                         number_particles=total, number_up=n_up, spin_inversion=None, symmetries=None)
