"""The reference's golden-file test programs and benchmark harnesses, on this library.

    python -m lattice_symmetries_b200.reference_tests states    --kHamiltonian model.yaml --kRepresentatives golden.h5
    python -m lattice_symmetries_b200.reference_tests matvec    --kHamiltonian model.yaml --kVectors golden.h5
    python -m lattice_symmetries_b200.reference_tests benchmark --kHamiltonian model.yaml [--kRandomSeed 42]

``states``     chapel/test/TestStatesEnumeration.chpl: build the basis, compare with ``/representatives`` of the golden
               file element for element (``==`` on uint64, :27), print up to ten mismatches and the elapsed time.
``matvec``     chapel/test/TestMatrixVectorProduct.chpl: y = H x for ``/x`` of the golden file against ``/y`` with
               ``|a - b| <= max(kAbsTol, kRelTol * max(|a|, |b|))`` (:15-20; defaults 1e-13, 1e-11).
``benchmark``  chapel/benchmark/BenchmarkStatesEnumeration.chpl + BenchmarkMatrixVectorProduct.chpl: time the build and
               one product on a seeded random vector (kRandomSeed = 42); the reference prints elapsed seconds only,
               here states/s and matrix-elements/s as well.

The functions take the model as a YAML path and build / multiply with this package; ``backend`` swaps that for a
stand-in (``backend(parsed) -> (states, matvec, count_matrix_elements, build seconds, keep-alive)``), which is how the CPU
tests run them on the oracle.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

__all__ = ["Outcome", "check_states", "check_matvec", "benchmark", "approx_equal", "main"]


@dataclass
class Outcome:
    ok: bool
    elapsed: float
    mismatches: List[str] = field(default_factory=list)
    details: dict = field(default_factory=dict)


def approx_equal(a, b, atol: float = 1e-13, rtol: float = 1e-11):
    """chapel/test/TestMatrixVectorProduct.chpl:15-20, element-wise."""
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b) <= np.maximum(atol, rtol * np.maximum(np.abs(a), np.abs(b)))


def _device_backend(parsed):
    """Build the basis and the operator on the GPU(s) with this package."""
    from .operator import Operator
    basis = parsed.model.basis()
    t0 = time.perf_counter()
    basis.build()
    build_s = time.perf_counter() - t0
    op = Operator(basis, parsed.hamiltonian) if parsed.hamiltonian is not None else None
    count = (lambda: op.count_matrix_elements()) if op is not None else (lambda: 0)
    matvec = (lambda x: op.apply_to_state_vector(np.ascontiguousarray(x, dtype=np.float64))) if op is not None else None
    return np.asarray(basis.states), matvec, count, build_s, (basis, op)


def _prepare(model_file, backend):
    from .config import parse_yaml_file
    parsed = parse_yaml_file(model_file)
    made = (backend or _device_backend)(parsed)   # (states, matvec or None, count_matrix_elements, build seconds, keep-alive)
    return (parsed,) + tuple(made[:4])


def check_states(model_file, representatives_file, dataset: str = "/representatives", backend: Optional[Callable] = None,
                 log: Callable[[str], None] = print) -> Outcome:
    from . import hdf5
    t0 = time.perf_counter()
    parsed, states, _, _, build_s = _prepare(model_file, backend)
    elapsed = build_s or (time.perf_counter() - t0)
    reference = hdf5.read_dataset(representatives_file, dataset)
    mismatches: List[str] = []
    if reference.shape != states.shape:
        mismatches.append(f"number of representatives: {states.shape[0]} (computed) != {reference.shape[0]} (expected)")
    n = min(reference.shape[0], states.shape[0])
    bad = np.nonzero(reference[:n] != states[:n])[0]
    for i in bad[:10]:                                       # TestStatesEnumeration.chpl:28-36
        mismatches.append(f"at index {int(i)}: {int(reference[i])} != {int(states[i])}")
    for line in mismatches:
        log(line)
    log(f"{elapsed}")
    return Outcome(not mismatches, elapsed, mismatches, {"dim": int(states.shape[0])})


def check_matvec(model_file, vectors_file, x: str = "/x", y: str = "/y", atol: float = 1e-13, rtol: float = 1e-11,
                 backend: Optional[Callable] = None, log: Callable[[str], None] = print) -> Outcome:
    from . import hdf5
    parsed, states, matvec, _, _ = _prepare(model_file, backend)
    if matvec is None:
        raise ValueError(f"{model_file}: the configuration has no 'hamiltonian'")
    xs = hdf5.read_dataset(vectors_file, x)
    ys = hdf5.read_dataset(vectors_file, y)
    xs, ys = (xs[0], ys[0]) if xs.ndim == 2 else (xs, ys)     # f64[1, dim] in the reference's files (:7-11)
    if xs.shape[0] != states.shape[0]:
        raise ValueError(f"{vectors_file}: vectors of length {xs.shape[0]} on a basis of {states.shape[0]} states")
    t0 = time.perf_counter()
    zs = matvec(xs)
    elapsed = time.perf_counter() - t0
    close = approx_equal(zs, ys, atol, rtol)
    log(str(bool(close.all())).lower())
    mismatches = [f"at {int(i)}: {zs[i]!r} (computed) != {ys[i]!r} (expected); Δ = {abs(zs[i] - ys[i])!r}"
                  for i in np.nonzero(~close)[0][:10]]    # TestMatrixVectorProduct.chpl:47-57
    for line in mismatches:
        log(line)
    log(f"{elapsed}")
    return Outcome(bool(close.all()), elapsed, mismatches,
                   {"dim": int(states.shape[0]), "max_abs_err": float(np.max(np.abs(zs - ys))) if zs.size else 0.0})


def benchmark(model_file, seed: int = 42, run_matvec: bool = True, repeats: int = 3, backend: Optional[Callable] = None,
              log: Callable[[str], None] = print) -> Outcome:
    t0 = time.perf_counter()
    parsed, states, matvec, count, build_s = _prepare(model_file, backend)
    build_s = build_s or (time.perf_counter() - t0)
    dim = int(states.shape[0])
    log(f"Hilbert space dimension: {dim}")
    details = {"dim": dim, "build_s": build_s, "representatives_per_s": dim / build_s if build_s > 0 else float("inf")}
    elapsed = build_s
    if run_matvec and matvec is not None:
        x = np.random.default_rng(seed).random(dim)            # Random.fillRandom: uniform in [0, 1)
        times = []
        for _ in range(max(1, repeats)):
            t0 = time.perf_counter()
            matvec(x)
            times.append(time.perf_counter() - t0)
        elements = int(count()) + dim
        details.update(matvec_s=min(times), matrix_elements=elements, matrix_elements_per_s=elements / min(times))
        elapsed = min(times)
    log(f"{elapsed}")
    return Outcome(True, elapsed, [], details)


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    import json
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="program", required=True)
    s = sub.add_parser("states")
    s.add_argument("--kHamiltonian", required=True)
    s.add_argument("--kRepresentatives", required=True)
    m = sub.add_parser("matvec")
    m.add_argument("--kHamiltonian", required=True)
    m.add_argument("--kVectors", required=True)
    m.add_argument("--kAbsTol", type=float, default=1e-13)
    m.add_argument("--kRelTol", type=float, default=1e-11)
    b = sub.add_parser("benchmark")
    b.add_argument("--kHamiltonian", required=True)
    b.add_argument("--kRandomSeed", type=int, default=42)
    b.add_argument("--kRunMatrixVectorProduct", type=int, default=1)
    args = ap.parse_args(argv)
    if args.program == "states":
        out = check_states(args.kHamiltonian, args.kRepresentatives)
    elif args.program == "matvec":
        out = check_matvec(args.kHamiltonian, args.kVectors, atol=args.kAbsTol, rtol=args.kRelTol)
    else:
        out = benchmark(args.kHamiltonian, args.kRandomSeed, bool(args.kRunMatrixVectorProduct))
        print(json.dumps(out.details))
    return 0 if out.ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
