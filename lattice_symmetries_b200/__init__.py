"""B200-native hot paths of lattice-symmetries: symmetry-adapted basis
construction and the sparse Hamiltonian matvec, as CUDA kernels for sm_100a
behind the reference's ``ls_hs_*`` / ``ls_chpl_*`` C ABI.

The Python layer mirrors the reference's user API
(python/lattice_symmetries/__init__.py): ``Symmetry``, ``Symmetries``,
``SpinBasis``, ``SpinlessFermionBasis``, ``SpinfulFermionBasis``, ``Expr``,
``Operator``.  It is host-side set-up only; all compute goes through
``liblattice_symmetries_b200.so`` and fails loudly when the library or a CUDA
device is missing.
"""
from .symmetry import Symmetry, Symmetries
from .expr import Expr, NonbranchingTerm, compile_terms
from ._lib import lib, LIB_PATH
from .basis import Basis, SpinBasis, SpinlessFermionBasis, SpinfulFermionBasis
from .operator import Operator
from . import lattices
from .config import load_yaml_config

__all__ = [
    "Symmetry", "Symmetries", "Expr", "NonbranchingTerm", "compile_terms", "Basis", "SpinBasis",
    "SpinlessFermionBasis", "SpinfulFermionBasis", "Operator", "lattices", "lib", "LIB_PATH", "load_yaml_config",
]
