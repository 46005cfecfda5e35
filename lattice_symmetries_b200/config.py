"""The reference's YAML model files -> inputs of the two hot paths (host-side set-up).

Every model of the reference -- its tests, its benchmarks, BASELINE.json's configs[0]
(``chapel/data/heisenberg_chain_24_symm.yaml``) -- is a YAML file read by the Haskell host
(``ls_hs_load_yaml_config``: haskell/src/LatticeSymmetries/Yaml.hs:14-42, basis header
Basis.hs:270-319, symmetry Group.hs:70-73, term Expr.hs:374-384) and handed to the kernels as a basis plus
non-branching terms.  GHC is absent here, so this module reads the same files into the same objects:

    basis:        particle: spin-1/2 (default) | spinless-fermion | spinful-fermion
                  number_spins, hamming_weight?, spin_inversion?, symmetries?: [{permutation, sector}]
                  number_sites, number_particles?: N | [N_up, N_down]
    hamiltonian:  terms: [{expression, sites?, particle?}]   (summed; other keys -- name, lattice anchors -- are ignored)
    observables:  [ {terms: [...]}, ... ]

``load_yaml_config`` mirrors python/lattice_symmetries/__init__.py:762-772: it returns
``Config(basis, hamiltonian, observables)`` with a ``Basis`` and ``Operator`` s of this package (the reference raises
``NotImplementedError`` on observables; here they are loaded).  ``parse_config`` returns the plain description
(a :class:`lattices.Model`) without touching the library, which is what the CPU tests feed to the oracle.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Any, Dict, List, Optional

from .expr import Expr
from .lattices import Model
from .symmetry import Symmetries, Symmetry

__all__ = ["Config", "ParsedConfig", "parse_config", "parse_yaml_file", "load_yaml_config", "basis_header", "state_to_string"]

Config = namedtuple("Config", ["basis", "hamiltonian", "observables"], defaults=[None, None])
ParsedConfig = namedtuple("ParsedConfig", ["model", "hamiltonian", "observables", "extra"])

_PARTICLES = {"spin-1/2": "spin-1/2", "spin": "spin-1/2", "spinless-fermion": "spinless-fermion",
              "spinful-fermion": "spinful-fermion"}


def _particle(value, default: Optional[str] = "spin-1/2") -> Optional[str]:
    if value is None:
        return default
    if value not in _PARTICLES:
        raise ValueError(f"invalid particle type: {value!r}; expected one of {sorted(set(_PARTICLES.values()))}")
    return _PARTICLES[value]


def _int(value, what: str, optional: bool = False) -> Optional[int]:
    if value is None and optional:
        return None
    if isinstance(value, bool) or not isinstance(value, int):
        raise ValueError(f"{what}: expected an integer, got {value!r}")
    return int(value)


def _expression(spec: Dict[str, Any], particle: str) -> Expr:
    """Expr.hs:374-384 (``exprFromJSON``): ``expression`` with optional ``sites`` rows and an optional ``particle``
    that must agree with the basis."""
    if not isinstance(spec, dict) or "expression" not in spec:
        raise ValueError(f"a term needs an 'expression': {spec!r}")
    tp = _particle(spec.get("particle"), None)
    if tp is not None and tp != particle:
        raise ValueError(f"invalid particle type: {tp}; expected {particle}")
    return Expr(str(spec["expression"]), sites=spec.get("sites"))


def _operator(spec: Dict[str, Any], particle: str) -> Expr:
    """Yaml.hs:38-42 (``operatorFromJSON``): the non-empty list ``terms``, summed."""
    terms = spec.get("terms") if isinstance(spec, dict) else None
    if not isinstance(terms, list) or not terms:
        raise ValueError("an operator needs a non-empty list of 'terms'")
    total = _expression(terms[0], particle)
    for t in terms[1:]:
        total = total + _expression(t, particle)
    return total


def parse_config(config: Dict[str, Any], name: str = "config") -> ParsedConfig:
    """A decoded YAML / JSON document -> (model without an expression if there is no hamiltonian, hamiltonian Expr or
    None, observable Exprs, the keys this layer does not interpret)."""
    if not isinstance(config, dict) or not isinstance(config.get("basis"), dict):
        raise ValueError("the configuration needs a 'basis' mapping")
    b = config["basis"]
    particle = _particle(b.get("particle"))
    model_kwargs: Dict[str, Any] = {}
    if particle == "spin-1/2":
        n = _int(b.get("number_spins"), "basis.number_spins")
        hw = _int(b.get("hamming_weight"), "basis.hamming_weight", optional=True)
        inv = _int(b.get("spin_inversion"), "basis.spin_inversion", optional=True)
        if inv not in (None, 1, -1):
            raise ValueError(f"invalid spin_inversion: {inv}; expected 1, -1 or null")
        if hw is not None and not 0 <= hw <= n:
            raise ValueError(f"invalid hamming_weight: {hw}")
        gens = []
        for s in b.get("symmetries") or []:
            if not isinstance(s, dict) or "permutation" not in s or "sector" not in s:
                raise ValueError(f"a symmetry needs 'permutation' and 'sector': {s!r}")
            perm = [_int(i, "permutation entry") for i in s["permutation"]]
            if len(perm) != n:
                raise ValueError(f"permutation of length {len(perm)} on {n} spins")
            gens.append(Symmetry(perm, _int(s["sector"], "sector")))
        model_kwargs = dict(number_sites=n, hamming_weight=hw, spin_inversion=inv,
                            symmetries=Symmetries(gens) if gens else None, particle=particle)
    else:
        n = _int(b.get("number_sites"), "basis.number_sites")
        occ = b.get("number_particles")
        if particle == "spinful-fermion" and isinstance(occ, (list, tuple)):
            if len(occ) != 2:
                raise ValueError("number_particles: expected N or [N_up, N_down]")
            occ = (_int(occ[0], "number_particles[0]"), _int(occ[1], "number_particles[1]"))
        elif occ is not None:
            occ = _int(occ, "basis.number_particles")
        model_kwargs = dict(number_sites=n, particle=particle, number_particles=occ)
    hamiltonian = _operator(config["hamiltonian"], particle) if config.get("hamiltonian") is not None else None
    observables: List[Expr] = []
    if config.get("observables") is not None:
        if not isinstance(config["observables"], list):
            raise ValueError("'observables' must be a list of operators")
        observables = [_operator(o, particle) for o in config["observables"]]
    model = Model(name=name, expression=hamiltonian if hamiltonian is not None else Expr(""), **model_kwargs)
    extra = {k: v for k, v in config.items() if k not in ("basis", "hamiltonian", "observables")}
    return ParsedConfig(model, hamiltonian, observables, extra)


def parse_yaml_file(filename) -> ParsedConfig:
    import yaml
    from pathlib import Path
    with open(filename, "r", encoding="utf-8") as f:
        return parse_config(yaml.safe_load(f), name=Path(filename).stem)


def load_yaml_config(filename: str) -> Config:
    """python/lattice_symmetries/__init__.py:762-772: ``Config(basis, hamiltonian, observables)``."""
    from .operator import Operator
    parsed = parse_yaml_file(filename)
    basis = parsed.model.basis()
    hamiltonian = Operator(basis, parsed.hamiltonian) if parsed.hamiltonian is not None else None
    observables = [Operator(basis, e) for e in parsed.observables]
    return Config(basis, hamiltonian, observables)


def basis_header(particle: str, number_sites: int, hamming_weight=None, spin_inversion=None, symmetries=None,
                 number_particles=None) -> Dict[str, Any]:
    """The JSON object of a basis (``basisHeaderToJSON``, Basis.hs:289-302) -- also the ``basis`` mapping of a model
    file, so ``parse_config({"basis": basis_header(...)})`` gives the basis back."""
    if particle == "spin-1/2":
        return {"particle": "spin-1/2", "number_spins": int(number_sites), "hamming_weight": hamming_weight,
                "spin_inversion": spin_inversion, "symmetries": symmetries.json_object() if symmetries is not None else []}
    out: Dict[str, Any] = {"particle": particle, "number_sites": int(number_sites)}
    if particle == "spinful-fermion":
        if isinstance(number_particles, (tuple, list)):
            out["number_particles"] = [int(number_particles[0]), int(number_particles[1])]
        elif number_particles is not None:
            out["number_particles"] = int(number_particles)
    else:
        out["number_particles"] = None if number_particles is None else int(number_particles)
    return out


def state_to_string(state: int, number_bits: int, spinful: bool = False) -> str:
    """Pretty-printed basis state (Basis.hs:103-138): most significant bit first; spinful fermions as two kets, the
    upper half of the bits (the down sites, Basis.hs:621-628) first."""
    def bits(n: int, x: int) -> str:
        return "".join("1" if (x >> i) & 1 else "0" for i in reversed(range(n)))
    state = int(state)
    if spinful:
        half = number_bits // 2
        return "|" + bits(half, state >> half) + "⟩|" + bits(half, state) + "⟩"
    return "|" + bits(number_bits, state) + "⟩"
