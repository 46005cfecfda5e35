"""The bit-sliced orbit arithmetic of the CUDA kernels on the CPU: ``csrc/bitslice.cuh`` is ``__host__ __device__``, and
``tests/bitslice_harness.cpp`` compiles THAT header with g++ and runs its primitives through the group loops of
``csrc/orbit.cu`` (state_info) and ``csrc/basis_build.cu`` (is_representative).  Checked here against the oracle --
representatives, the minimising element's character, the alive / stabiliser-event bits -- for plane counts from 10 to
64, real and complex sectors, with and without spin inversion.  (The kernels themselves run in the ``-m gpu`` tier; this
pins the arithmetic idea they share -- planes renamed instead of bits permuted, min(y, ~y) = y ^ top(y), least
significant plane first -- where no GPU is needed.)"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import helpers as H

ROOT = Path(__file__).resolve().parent.parent
SOURCE = ROOT / "tests" / "bitslice_harness.cpp"
LIBRARY = ROOT / "tests" / "_build" / "libbitslice_harness.so"


@pytest.fixture(scope="module")
def harness():
    header = ROOT / "lattice_symmetries_b200" / "csrc" / "bitslice.cuh"
    if not LIBRARY.exists() or LIBRARY.stat().st_mtime < max(SOURCE.stat().st_mtime, header.stat().st_mtime):
        LIBRARY.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", str(SOURCE), "-o", str(LIBRARY)],
                       check=True)
    lib = C.CDLL(str(LIBRARY))
    lib.bitslice_selftest.restype = C.c_int
    lib.bitslice_selftest.argtypes = [C.c_uint64]
    lib.bitslice_state_info.restype = None
    lib.bitslice_state_info.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    lib.bitslice_is_representative.restype = C.c_int64
    lib.bitslice_is_representative.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int]
    return lib


def test_transposes(harness):
    for seed in range(5):
        assert harness.bitslice_selftest(seed) == 0


def _chain(n, hw, inv, t=0, p=0):
    from lattice_symmetries_b200 import lattices as L
    return lambda: (n, hw, inv, L.chain_symmetries(n, t, p))


def _kagome12():
    k12 = H.kagome12_complex_sector()
    return k12.number_sites, k12.hamming_weight, k12.spin_inversion, k12.symmetries


def _kagome(n, inv):
    def make():
        from lattice_symmetries_b200 import lattices as L
        return n, n // 2, inv, L.kagome_heisenberg(n, spin_inversion=inv).symmetries
    return make


CASES = {   # name -> () -> (sites, hamming weight, spin inversion, symmetries)
    "chain10_getting_started": lambda: (10, 5, -1, H.chain10_getting_started().symmetries),
    "chain24_inv": _chain(24, 12, 1),
    "chain24_noinv": _chain(24, 12, None),
    "chain20_k3_complex": _chain(20, 10, None, 3, None),
    "kagome12_complex": _kagome12,
    "kagome18": _kagome(18, None),
    "kagome24_inv": _kagome(24, 1),
    "chain32_inv_minus": _chain(32, 16, -1),
    "chain40_hw3": _chain(40, 3, None),
    "chain56_hw3": _chain(56, 3, None),
    "chain64_hw2": _chain(64, 2, None),
    "chain64_half": _chain(64, 32, 1),
}


@pytest.mark.parametrize("name", list(CASES))
def test_bitsliced_orbit_arithmetic_matches_the_oracle(harness, oracle, name):
    n, hw, inv, syms = CASES[name]()
    group = oracle.Group.from_symmetries(syms, n, inv)
    perm = np.ascontiguousarray(syms.permutations(), dtype=np.int32)
    G = perm.shape[0]
    assert perm.shape == (G, n) and G == len(group.re)
    rng = np.random.default_rng(42)
    words = 64
    states = H.random_fixed_hamming_states(rng, n, hw, 32 * words - 64)
    # some representatives, so that alive lanes and stabiliser events occur; and an ascending run (the low-16 shortcut)
    reps = group.state_info(states)[0][:32]
    low = min(16, n)
    run = np.sort(np.uint64(int(states[0]) & ~((1 << low) - 1)) | rng.choice(1 << low, size=32, replace=False).astype(np.uint64))
    states = np.ascontiguousarray(np.concatenate([states, reps, run]), dtype=np.uint64)
    assert states.shape[0] == 32 * words

    # ---- state_info ----
    got = np.zeros_like(states)
    arg = np.zeros(states.shape[0], dtype=np.int32)
    flip = np.zeros(states.shape[0], dtype=np.uint8)
    harness.bitslice_state_info(n, G, perm.ctypes.data, int(inv or 0), words, states.ctypes.data, got.ctypes.data,
                                arg.ctypes.data, flip.ctypes.data)
    want, chars, norms = group.state_info(states)
    assert np.array_equal(got, want)
    chi = np.where(arg >= 0, group.re[np.maximum(arg, 0)] + 1j * group.im[np.maximum(arg, 0)], 1.0 + 0j)
    chi = np.where(flip == 1, chi * float(inv or 1), chi)
    assert np.array_equal(chi.real, chars.real) and np.array_equal(chi.imag, chars.imag)
    assert np.all(flip[arg < 0] == 0) and (inv is not None or not flip.any())

    # ---- is_representative ----
    identity = int(np.nonzero(np.all(perm == np.arange(n), axis=1))[0][0])
    for low16 in (0, 1):
        alive = np.zeros(states.shape[0], dtype=np.uint8)
        events = np.zeros(states.shape[0], dtype=np.uint8)
        bad = harness.bitslice_is_representative(n, G, perm.ctypes.data, int(inv or 0), identity, words,
                                                 states.ctypes.data, alive.ctypes.data, events.ctypes.data, low16)
        assert bad == 0
        flags, sums = group.is_representative(states)
        assert np.array_equal(alive == 1, want == states)              # alive <=> the state is its own orbit minimum
        assert np.all(alive[flags == 1] == 1)                           # the oracle's representatives are alive ...
        quiet = (alive == 1) & (events == 0)                            # ... alive with a trivial stabiliser: norm^2 = 1/|G|
        assert np.all(flags[quiet] == 1) and np.all(sums[quiet] == 1.0)
        dead = (alive == 1) & (flags == 0)                              # alive but rejected: the character sum vanished,
        assert np.all(events[dead] == 1)                                # which takes a non-trivial stabiliser
        assert np.all(norms[dead] < 1e-6)
    assert alive.sum() >= 32
