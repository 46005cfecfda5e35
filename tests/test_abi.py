"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/lattice_symmetries_b200.h declares (and the Python
binding's prototype table covers exactly that set), struct layouts match the
reference header, and the product fails loudly -- no CPU fallback -- when no
CUDA device is present.  No compute calls here."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "lattice_symmetries_b200.h"


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b(ls_[a-z0-9_]+)\s*\(", text):
        name = m.group(1)
        # skip function-pointer typedef names and struct tags
        if name.endswith("_type") or name in ("ls_hs_scalar",):
            continue
        names.add(name)
    return sorted(names)


def test_header_declares_the_reference_surface():
    names = declared_functions()
    # symbols the Haskell host / Chapel / Python callers import
    for required in [
        "ls_internal_create_halide_kernel_data", "ls_internal_destroy_halide_kernel_data",
        "ls_hs_is_representative_halide_kernel", "ls_hs_state_info_halide_kernel",
        "ls_hs_create_state_index_binary_search_kernel_data", "ls_hs_destroy_state_index_binary_search_kernel_data",
        "ls_hs_state_index_binary_search_kernel", "ls_hs_build_representatives",
        "ls_hs_unchecked_set_representatives", "ls_hs_internal_destroy_external_array", "ls_hs_error",
        "ls_hs_fatal_error", "ls_hs_set_exception_handler", "ls_hs_state_index", "ls_hs_is_representative",
        "ls_hs_state_info", "ls_hs_internal_get_chpl_kernels", "ls_hs_internal_set_chpl_kernels",
        "ls_internal_operator_apply_diag_x1", "ls_internal_operator_apply_off_diag_x1",
        "ls_hs_internal_read_refcount", "ls_hs_internal_write_refcount", "ls_hs_internal_inc_refcount",
        "ls_hs_internal_dec_refcount", "ls_chpl_init", "ls_chpl_finalize", "ls_chpl_enumerate_representatives",
        "ls_chpl_operator_apply_diag", "ls_chpl_operator_apply_off_diag", "ls_chpl_matrix_vector_product",
    ]:
        assert required in names, required


def test_library_exports_every_declared_symbol():
    lib_path = ROOT / "lattice_symmetries_b200" / "liblattice_symmetries_b200.so"
    assert lib_path.exists(), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = C.CDLL(str(lib_path))
    missing = [n for n in declared_functions() if not hasattr(handle, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from lattice_symmetries_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()


def test_struct_layouts_match_reference_header():
    """kernels/lattice_symmetries_types.h:109-161 on x86-64 (static_asserts in
    csrc/abi_layout.cu pin the same numbers on the C++ side)."""
    from lattice_symmetries_b200 import _lib
    assert C.sizeof(_lib.chpl_external_array) == 24
    assert C.sizeof(_lib.ls_hs_scalar) == 16
    assert C.sizeof(_lib.ls_hs_basis_kernels) == 48
    g = _lib.ls_hs_permutation_group
    assert (g.masks.offset, g.shifts.offset, g.eigvals_re.offset, g.eigvals_im.offset, g.haskell_payload.offset) == \
        (16, 24, 32, 40, 48)
    b = _lib.ls_hs_basis
    assert (b.number_sites.offset, b.number_particles.offset, b.number_up.offset, b.particle_type.offset,
            b.spin_inversion.offset, b.state_index_is_identity.offset, b.requires_projection.offset,
            b.kernels.offset, b.representatives.offset, b.haskell_payload.offset) == \
        (4, 8, 12, 16, 20, 24, 25, 32, 40, 64)
    assert C.sizeof(b) == 72
    t = _lib.ls_hs_nonbranching_terms
    assert (t.v.offset, t.m.offset, t.l.offset, t.r.offset, t.x.offset, t.s.offset) == (8, 16, 24, 32, 40, 48)
    o = _lib.ls_hs_operator
    assert (o.basis.offset, o.off_diag_terms.offset, o.diag_terms.offset, o.haskell_payload.offset) == (8, 16, 24, 32)
    assert C.sizeof(_lib.ls_chpl_kernels) == 32


REFERENCE_KERNELS = Path("/root/reference/kernels")
HARNESS = ROOT / "tests" / "_build" / "abi_harness"


def build_harness() -> Path:
    """tests/abi_harness.c against the REFERENCE's own header (not ours), linked with the drop-in library.  The binary
    stays under tests/_build/ (git-ignored, travels to the GPU box, where /root/reference does not exist)."""
    import subprocess
    HARNESS.parent.mkdir(exist_ok=True)
    subprocess.run(
        ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", f"-I{REFERENCE_KERNELS}", "-DLS_NO_STD_COMPLEX",
         str(ROOT / "tests" / "abi_harness.c"), "-o", str(HARNESS), f"-L{ROOT / 'lattice_symmetries_b200'}",
         "-llattice_symmetries_b200", "-Wl,-rpath,$ORIGIN/../../lattice_symmetries_b200"], check=True)
    return HARNESS


@pytest.mark.skipif(not REFERENCE_KERNELS.is_dir(), reason="the reference tree is not mounted")
def test_reference_header_layout_matches_ours():
    """Every struct of kernels/lattice_symmetries_types.h as the reference's compiler lays it out (sizeof / offsetof
    printed by a C harness that includes THAT header) against the binding's structs, which mirror
    include/lattice_symmetries_b200.h field for field (and csrc/abi_layout.cu pins the same numbers in C++)."""
    import json
    import subprocess
    from lattice_symmetries_b200 import _lib
    harness = build_harness()
    layout = json.loads(subprocess.run([str(harness), "layout"], check=True, capture_output=True, text=True).stdout)
    checked = 0
    for key, value in layout.items():
        if key.startswith("sizeof "):
            assert C.sizeof(getattr(_lib, key.split()[1])) == value, key
            checked += 1
        elif "." in key:
            struct, field = key.split(".")
            assert getattr(getattr(_lib, struct), field).offset == value, key
            checked += 1
    assert checked >= 50
    assert (layout["LS_HS_SPIN"], layout["LS_HS_SPINFUL_FERMION"], layout["LS_HS_SPINLESS_FERMION"]) == (0, 1, 2)


def test_vtable_registration_needs_no_device():
    """ls_chpl_init_kernels (LatticeSymmetries.chpl:18-33) fills the four vtable slots."""
    from lattice_symmetries_b200 import _lib
    _lib.lib.ls_chpl_init_kernels()
    k = _lib.lib.ls_hs_internal_get_chpl_kernels().contents
    assert all([k.enumerate_states, k.operator_apply_off_diag, k.operator_apply_diag, k.matrix_vector_product])


def test_refcount_helpers():
    from lattice_symmetries_b200 import _lib
    v = C.c_int(1)
    assert _lib.lib.ls_hs_internal_inc_refcount(C.byref(v)) == 1 and v.value == 2
    assert _lib.lib.ls_hs_internal_dec_refcount(C.byref(v)) == 2 and v.value == 1
    _lib.lib.ls_hs_internal_write_refcount(C.byref(v), 7)
    assert _lib.lib.ls_hs_internal_read_refcount(C.byref(v)) == 7


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail through
    ls_hs_error (-> RuntimeError), never compute on the host."""
    from lattice_symmetries_b200 import _lib
    if _lib.lib.ls_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import lattice_symmetries_b200 as ls
    _lib._initialised = False
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ls.SpinBasis(4)
    _lib._initialised = False


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "lattice_symmetries_b200"
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".hpp"):
            text = path.read_text()
            assert not re.search(r"(from|import)\s+oracle|liboracle|ls_oracle|#include[^\n]*oracle|_ref/libref", text), path
