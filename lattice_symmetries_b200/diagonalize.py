"""Exact diagonalisation driver: YAML model in, HDF5 results out -- the reference's ``Diagonalize`` program on top of
the two hot paths (chapel/src/Diagonalize.chpl:166-177 options, :227-256 file layout, :258-333 main).

    python -m lattice_symmetries_b200.diagonalize --input model.yaml --kOutput out.h5 --numEvals 2 --kEps 1e-8
    torchrun --nproc-per-node 8 ... -m lattice_symmetries_b200.diagonalize --input kagome36.yaml

What it does, in the reference's order: read the model (``config.parse_yaml_file`` = ``loadConfigFromYaml``); create
the output file with the groups ``basis``, ``hamiltonian`` (``makeFile`` / ``makeGroup``); build the representatives
on the GPU(s) -- or, on one GPU, reuse ``basis/representatives`` if the output file already holds them
(``makeBasisStates`` :227-245) -- and store them; find the ``numEvals`` lowest eigenpairs (the reference calls PRIMME,
here the on-device thick-restart Lanczos of ``lanczos.py``; ``kEps`` is the residual tolerance relative to
max(1, |E|), ``kMaxBasisSize`` the Krylov basis kept in HBM, ``kMaxBlockSize`` > 1 the block method over the library's
block product); write ``hamiltonian/eigenvectors`` f64[numEvals, dim],
``hamiltonian/eigenvalues`` and ``hamiltonian/residuals`` f64[numEvals] (``saveEigenvectors`` :247-256).  With several
ranks every rank writes its own rows of the one file (``hdf5.write_rows``); the sorted contiguous ranges in rank
order are the reference's block layout.

The solver and the file logic are host code over the small operator interface of ``lanczos._wrap``; the GPU enters
through ``operator_factory`` (default: build the basis and the operator with this package), which is how the CPU
tests drive the whole program against the oracle.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, List, Optional

import numpy as np

__all__ = ["DiagonalizeResult", "diagonalize", "run_options", "main"]


@dataclass
class DiagonalizeResult:
    eigenvalues: np.ndarray
    residuals: np.ndarray
    dim: int
    output: Path
    matvecs: int
    converged: bool
    reused_representatives: bool = False
    seconds: dict = field(default_factory=dict)


def _device_operator(parsed, cached_representatives: Optional[np.ndarray] = None):
    """The default ``operator_factory(parsed, cached_representatives or None)``: this package's Basis + Operator on the
    GPU(s) -> (solver interface, local block of the representatives, objects to keep alive)."""
    from .lanczos import _wrap
    from .operator import Operator
    basis = parsed.model.basis()
    if cached_representatives is not None:
        basis.unchecked_set_representatives(cached_representatives)
    else:
        basis.build()
    op = Operator(basis, parsed.hamiltonian)
    sh = _wrap(op)
    return sh, np.asarray(basis.states), (basis, op)


def _is_real(parsed) -> bool:
    """Is the matrix real in the symmetry-adapted basis: real characters and real couplings."""
    from .expr import compile_terms
    m = parsed.model
    # (exact test on the rational phases: sin(-2 pi / 2) is -1.2e-16 in floating point, not 0)
    if m.symmetries is not None and any((2 * g.phase).denominator != 1 for g in m.symmetries.elements):
        return False
    return all(complex(t.v).imag == 0.0 for t in compile_terms(parsed.hamiltonian, m.number_sites))


def _complex_dtype():
    import torch
    return torch.complex128


def diagonalize(input, output="exact_diagonalization_output.h5", num_evals: int = 1, eps: float = 1e-6,
                max_basis_size: int = 0, max_restarts: int = 200, seed: int = 42, max_block_size: int = 1,
                operator_factory: Optional[Callable] = None, barrier: Optional[Callable[[], None]] = None,
                log: Optional[Callable[[str], None]] = None) -> DiagonalizeResult:
    """Run the program described in the module docstring; ``barrier()`` must synchronise the ranks when there are
    several (``torch.distributed.barrier``).  Returns the eigenvalues / residuals (identical on every rank)."""
    from . import hdf5
    from .config import parse_yaml_file
    from .lanczos import lanczos_block_thick_restart, lanczos_thick_restart
    say = log or (lambda s: None)
    if num_evals < 1:
        raise ValueError(f"invalid numEvals: {num_evals}")
    parsed = parse_yaml_file(input) if not hasattr(input, "model") else input
    if parsed.hamiltonian is None:
        raise ValueError(f"{input}: the configuration has no 'hamiltonian'")
    output = Path(output)
    factory = operator_factory or _device_operator
    seconds = {}

    # representatives: from the output file when it has them (one rank), else built here
    cached = None
    if output.exists():
        try:
            with hdf5.File(output) as f:
                if f.exists("basis/representatives"):
                    cached = f.read("basis/representatives")
        except hdf5.Hdf5Error:
            cached = None
    if barrier is not None:
        cached = None   # several ranks: always build (the sharded build is the fast path; a cache would have to be re-split)
    t0 = time.perf_counter()
    made = factory(parsed, cached)
    sh, local_reps = made[0], made[1]
    L = sh.layout
    if cached is not None and (L.world != 1 or cached.shape[0] != L.dim):
        raise ValueError(f"{output}: cached basis/representatives do not match this run")
    seconds["basis"] = time.perf_counter() - t0
    say(f"Hilbert space dimension: {L.dim}" + (" (representatives read from the output file)" if cached is not None else ""))
    k = min(int(num_evals), L.dim)
    if k < 1:
        raise ValueError("the basis is empty")

    # the output file: created once at its final size, then filled block by block.  (The reference is real-only,
    # DistributedMatrixVector.chpl:1090-1091; a complex sector or complex couplings give complex128 eigenvectors,
    # stored as the {r, i} compound h5py reads as complex.)
    vec_dtype = np.float64 if _is_real(parsed) else np.complex128
    sync = barrier or (lambda: None)
    if L.world > 1 and barrier is None:
        raise ValueError("several ranks need a barrier")
    if L.rank == 0:
        hdf5.create(output, {"basis/representatives": hdf5.DatasetSpec((L.dim,), np.uint64),
                             "hamiltonian/eigenvectors": hdf5.DatasetSpec((k, L.dim), vec_dtype),
                             "hamiltonian/eigenvalues": hdf5.DatasetSpec((k,), np.float64),
                             "hamiltonian/residuals": hdf5.DatasetSpec((k,), np.float64)})
    sync()
    hdf5.write_rows(output, "basis/representatives", np.asarray(local_reps, dtype=np.uint64), L.row_begin)

    # eigenpairs
    t0 = time.perf_counter()
    m = int(max_basis_size) if max_basis_size else max(2 * k + 16, 24)
    m = min(max(m, k + 2), L.dim)   # room for a restart, never more vectors than the space has dimensions
    dtype = None if _is_real(parsed) else _complex_dtype()
    if max_block_size > 1:   # kMaxBlockSize: several vectors per pass over the matrix elements (block product)
        bsz = min(int(max_block_size), L.dim)
        res = lanczos_block_thick_restart(sh, k=k, block_size=bsz, basis_size=min(max(m, k + 2 * bsz), L.dim), tol=eps,
                                          max_restarts=max_restarts, seed=seed, dtype=dtype)
    else:
        res = lanczos_thick_restart(sh, k=k, basis_size=m, tol=eps, max_restarts=max_restarts, seed=seed, dtype=dtype)
    seconds["eigensolver"] = time.perf_counter() - t0
    evals = np.asarray(res.energies, dtype=np.float64)
    resid = np.asarray(res.residuals, dtype=np.float64)
    say(f"Obtained eigenvalues: {evals.tolist()}")
    say(f"Residual norms:       {resid.tolist()}")

    vecs = res.eigenvectors
    vecs = vecs.detach().cpu().numpy() if hasattr(vecs, "detach") else np.asarray(vecs)
    hdf5.write_rows(output, "hamiltonian/eigenvectors", np.ascontiguousarray(vecs, dtype=vec_dtype), L.row_begin)
    if L.rank == 0:
        hdf5.write_rows(output, "hamiltonian/eigenvalues", evals, 0)
        hdf5.write_rows(output, "hamiltonian/residuals", resid, 0)
    sync()
    return DiagonalizeResult(evals, resid, L.dim, output, res.matvecs, bool(res.converged), cached is not None, seconds)


def run_options(input, output=None, num_evals=None, max_basis_size=None, max_block_size=None) -> dict:
    """The options of a run: what the caller gives wins, else what the model file itself says -- the reference's model
    files carry ``output``, ``number_vectors``, ``max_primme_basis_size``, ``max_primme_block_size``
    (chapel/data/heisenberg_kagome_16.yaml:16-18, heisenberg_square_6x6.yaml:72-77) -- else the driver's defaults
    (chapel/src/Diagonalize.chpl:166-177)."""
    from .config import parse_yaml_file
    extra = parse_yaml_file(input).extra

    def pick(given, key, default):
        return given if given is not None else extra.get(key, default)

    return {"output": str(pick(output, "output", "exact_diagonalization_output.h5")),
            "num_evals": int(pick(num_evals, "number_vectors", 1)),
            "max_basis_size": int(pick(max_basis_size, "max_primme_basis_size", 0)),
            "max_block_size": int(pick(max_block_size, "max_primme_block_size", 1))}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--input", default="data/heisenberg_chain_10.yaml")                # Diagonalize.chpl:166
    ap.add_argument("--kOutput", default=None)                                         # :167
    ap.add_argument("--numEvals", type=int, default=None)                              # :168
    ap.add_argument("--kEps", type=float, default=1e-6)                                # :169
    ap.add_argument("--kMaxBasisSize", type=int, default=None)                         # :172
    ap.add_argument("--kMaxBlockSize", type=int, default=None)                         # :175
    ap.add_argument("--maxRestarts", type=int, default=200)
    args = ap.parse_args(argv)
    options = run_options(args.input, args.kOutput, args.numEvals, args.kMaxBasisSize, args.kMaxBlockSize)

    import os
    barrier = None
    rank = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        from .distributed import init_communicator, init_process
        init_process()
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        init_communicator()
        barrier, rank = dist.barrier, dist.get_rank()
    log = (lambda s: print(s, flush=True)) if rank == 0 else None
    res = diagonalize(args.input, options["output"], options["num_evals"], args.kEps, options["max_basis_size"],
                      args.maxRestarts, max_block_size=options["max_block_size"], barrier=barrier, log=log)
    if rank == 0:
        print(f"{res.dim} states, {res.matvecs} products, converged: {res.converged}; "
              f"basis {res.seconds['basis']:.3f} s, eigensolver {res.seconds['eigensolver']:.3f} s -> {res.output}")
    return 0 if res.converged else 1


if __name__ == "__main__":
    raise SystemExit(main())
