"""Ground state by Lanczos iteration with all vectors resident in HBM.

The reference delegates the eigensolve to scipy's ``eigsh``
(python/example/getting_started.py:49) or PRIMME (chapel/src/Diagonalize.chpl:
134-162, 298-325); both bounce every Lanczos vector through the host.  Here the
three-term recurrence keeps its vectors on the GPU(s): the matvec is the
library's device entry point (``ls_b200_matvec_device``), the vector algebra is
torch on the same stream (plumbing), the global dots are one all-reduce of a
scalar, and only the tridiagonal coefficients reach the host.

No re-orthogonalisation: ghost copies do not disturb the extremal eigenvalue.
The eigenvector, when requested, is accumulated in a second pass that replays
the recurrence from the same start vector (two-pass Lanczos: 3 vectors of HBM
instead of one per iteration).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

__all__ = ["LanczosResult", "lanczos_ground_state"]


@dataclass
class LanczosResult:
    energy: float
    iterations: int
    residual: float
    converged: bool
    eigenvector: Optional[object] = None   # padded replicated torch tensor (ShardedOperator layout)
    alphas: Optional[np.ndarray] = None
    betas: Optional[np.ndarray] = None


def _lowest(alphas, betas):
    from scipy.linalg import eigh_tridiagonal
    if len(alphas) == 1:
        return float(alphas[0]), np.ones(1)
    w, v = eigh_tridiagonal(np.asarray(alphas), np.asarray(betas[: len(alphas) - 1]), select="i", select_range=(0, 0))
    return float(w[0]), v[:, 0]


def _start_vector(sharded, seed: int):
    """Deterministic start vector that does not depend on the number of ranks:
    every rank draws the full vector from the same seed (host), keeps it all
    (x is replicated anyway)."""
    import torch
    L = sharded.layout
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    v = sharded.empty_vector()
    chunk = 1 << 24
    for lo in range(0, L.dim, chunk):
        hi = min(L.dim, lo + chunk)
        v[lo:hi].copy_(torch.randn(hi - lo, dtype=torch.float64, generator=g))
    return v


def lanczos_ground_state(operator, max_iters: int = 300, tol: float = 1e-12, seed: int = 42,
                         compute_eigenvector: bool = False, group=None, check_every: int = 5) -> LanczosResult:
    """Lowest eigenvalue (and optionally eigenvector) of a real symmetric ``Operator``.

    ``tol`` bounds the Ritz residual |beta_k s_k| relative to |E0|.
    Works on one GPU or, under ``torch.distributed``, on the rank's row shard
    (see :class:`distributed.ShardedOperator`)."""
    import torch
    from . import _lib
    from .distributed import ShardedOperator, init_process

    init_process(torch.cuda.current_device() if torch.cuda.is_available() else None)
    sh = operator if isinstance(operator, ShardedOperator) else ShardedOperator(operator, group)

    def run(accumulate_with=None):
        v = _start_vector(sh, seed)
        nrm = torch.sqrt(sh.dot(v, v))
        v /= nrm
        v_prev = sh.empty_vector()
        w = sh.empty_vector()
        out = sh.empty_vector() if accumulate_with is not None else None
        alphas, betas = [], []
        beta = 0.0
        energy, resid, converged = float("nan"), float("inf"), False
        n_steps = len(accumulate_with) if accumulate_with is not None else max_iters
        for k in range(n_steps):
            if out is not None:
                out.add_(v, alpha=float(accumulate_with[k]))
            sh.matvec(v, w)
            alpha = float(sh.dot(v, w).item())
            alphas.append(alpha)
            w.add_(v, alpha=-alpha)
            if k > 0:
                w.add_(v_prev, alpha=-beta)
            beta = float(torch.sqrt(sh.dot(w, w)).item())
            betas.append(beta)
            if accumulate_with is None and ((k + 1) % check_every == 0 or k + 1 == n_steps or beta < 1e-14):
                energy, s = _lowest(alphas, betas)
                resid = abs(beta * s[-1])
                if resid <= tol * max(1.0, abs(energy)) or beta < 1e-14:
                    converged = True
                    break
            if beta < 1e-14:
                break
            v_prev, v, w = v, w, v_prev
            v /= beta
        _lib.lib.ls_b200_matvec_sync()
        _lib.check_error()
        return alphas, betas, energy, resid, converged, out

    alphas, betas, energy, resid, converged, _ = run()
    result = LanczosResult(energy, len(alphas), resid, converged, None, np.array(alphas), np.array(betas))
    if compute_eigenvector:
        _, s = _lowest(alphas, betas)
        *_, vec = run(accumulate_with=s)
        vec /= torch.sqrt(sh.dot(vec, vec))
        result.eigenvector = vec
    return result
