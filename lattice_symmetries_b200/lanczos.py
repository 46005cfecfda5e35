"""Eigensolvers with all vectors resident in HBM.

The reference delegates the eigensolve to scipy's ``eigsh`` (python/example/getting_started.py:49) or PRIMME
(chapel/src/Diagonalize.chpl:134-162, 166-177 ``numEvals``, 293-325); both bounce every Lanczos vector through
the host.  Here the vectors never leave the GPU(s):

* :func:`lanczos_ground_state` -- three-term recurrence on three local vectors (the only option for kagome-42,
  whose vectors are 9.6 GB per rank).  The recurrence coefficients stay ON THE DEVICE: alpha and beta are
  1-element tensors produced by a local dot + one in-place all-reduce on the library's communicator, consumed
  by ``addcmul_`` / ``div_`` without ever being read -- the host only looks at the tridiagonal matrix every
  ``check_every`` iterations, so the launch queue never drains.
* :func:`lanczos_thick_restart` -- k lowest eigenpairs with a bounded Krylov basis (m vectors in HBM), full
  re-orthogonalisation inside the basis and Wu-Simon thick restarts; real or complex128 vectors.

* :func:`lanczos_block_thick_restart` -- the same with a BLOCK of b vectors per step.  The integer work of a product
  (canonicalise + rank every matrix element) is per matrix element, not per vector, so the library's block product
  (``ls_b200_matvec_block_device``; kagome-36: 27.7 ms per vector at b = 8 against 80 ms alone) makes b right-hand
  sides cost far less than b products; a block method also resolves (near-)degenerate levels a single vector crawls
  through.  (``kMaxBlockSize`` of the reference's driver, chapel/src/Diagonalize.chpl:175.)

All run on one GPU (an ``Operator``) or on a rank's row block of a sharded basis
(:class:`distributed.DistributedOperator`); the matvec is the library's device entry point either way.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

__all__ = ["LanczosResult", "lanczos_ground_state", "lanczos_thick_restart", "lanczos_block_thick_restart"]


@dataclass
class LanczosResult:
    energy: float
    iterations: int
    residual: float
    converged: bool
    eigenvector: Optional[object] = None   # torch tensor: this rank's rows
    alphas: Optional[np.ndarray] = None
    betas: Optional[np.ndarray] = None
    energies: Optional[np.ndarray] = None  # thick restart: the k lowest eigenvalues
    residuals: Optional[np.ndarray] = None
    eigenvectors: Optional[object] = None  # thick restart: [k, local rows]
    matvecs: int = 0


class _SingleDevice:
    """The DistributedOperator interface over a plain Operator on one GPU."""

    device = "cuda"

    def __init__(self, operator):
        from .distributed import Layout
        self.op = operator
        dim = operator.basis.number_states
        self.layout = Layout(1, 0, dim, 0, dim, 0, 0, 0, [0, dim])

    def empty_vector(self, dtype=None):
        import torch
        return torch.zeros(self.layout.dim, dtype=dtype or torch.float64, device="cuda")

    def matvec(self, x, y, mode=None):
        import torch
        self.op.matvec_device(x.data_ptr(), y.data_ptr(), complex_vectors=x.dtype == torch.complex128)

    def matvec_block(self, X, Y):
        """Y[v] = H X[v] for the rows of two contiguous [b, dim] float64 tensors: ONE pass over the matrix elements."""
        import torch
        assert X.dtype == torch.float64 and X.is_contiguous() and Y.is_contiguous() and X.shape == Y.shape
        self.op.matvec_block_device(X.shape[0], X.data_ptr(), X.shape[1], Y.data_ptr(), Y.shape[1])

    def sync(self):
        from . import _lib
        _lib.lib.ls_b200_matvec_sync()
        _lib.check_error()

    def dot(self, a, b):
        import torch
        return torch.vdot(a, b).reshape(1)


def _wrap(operator):
    """An ``Operator`` (one GPU, or a rank's block of a sharded basis) behind the small interface the solvers use:
    ``layout``, ``device``, ``empty_vector``, ``matvec``, ``dot``, ``sync``.  Anything that already has it -- a
    :class:`DistributedOperator`, or a stand-in on the CPU in the tests -- passes through."""
    from .distributed import DistributedOperator, init_process, layout_of
    import torch
    if hasattr(operator, "layout") and hasattr(operator, "matvec") and not isinstance(operator, DistributedOperator):
        return operator   # a stand-in that brings its own device (tests)
    # torch's vector algebra and the library's kernels / collectives must share one stream
    init_process(torch.cuda.current_device() if torch.cuda.is_available() else None)
    if isinstance(operator, DistributedOperator):
        return operator
    try:
        layout_of(operator.basis)
    except ValueError:
        return _SingleDevice(operator)
    return DistributedOperator(operator)


def _lowest(alphas, betas):
    from scipy.linalg import eigh_tridiagonal
    if len(alphas) == 1:
        return float(alphas[0]), np.ones(1)
    w, v = eigh_tridiagonal(np.asarray(alphas), np.asarray(betas[: len(alphas) - 1]), select="i", select_range=(0, 0))
    return float(w[0]), v[:, 0]


def _start_vector(sh, seed: int, dtype=None):
    """Deterministic, independent of the number of ranks: entry i depends on (seed, global row i) only."""
    import torch
    from .distributed import hashed_vector
    L = sh.layout
    device = getattr(sh, "device", "cuda")
    v = hashed_vector(L.row_begin, L.row_end, seed, device=device)
    if dtype == torch.complex128:
        v = torch.complex(v, hashed_vector(L.row_begin, L.row_end, seed + 1000003, device=device))
    return v


def lanczos_ground_state(operator, max_iters: int = 300, tol: float = 1e-12, seed: int = 42,
                         compute_eigenvector: bool = False, check_every: int = 10,
                         time_limit_s: Optional[float] = None, progress=None,
                         energy_tol: Optional[float] = None, checkpoint: Optional[str] = None,
                         checkpoint_every: int = 0, resume: bool = False) -> LanczosResult:
    """Lowest eigenvalue (and optionally eigenvector) of a real symmetric ``Operator``.

    ``tol`` bounds the Ritz residual |beta_k s_k| relative to |E0|.  No re-orthogonalisation: ghost copies do not
    disturb the extremal eigenvalue.  The eigenvector, when requested, is accumulated in a second pass that
    replays the recurrence (two-pass Lanczos: 3 vectors of HBM instead of one per iteration).
    ``time_limit_s``: stop at the next check once this much wall time has passed (the result says whether it
    had converged).  ``progress(k, energy, residual)`` is called at every check.  ``energy_tol``: also stop when the
    Ritz value moved by less than ``energy_tol * |E0|`` over each of the last two checks (the eigenVALUE converges
    with the square of the residual, long before the residual itself is small).

    ``checkpoint`` (a path prefix) with ``checkpoint_every`` = n: every n iterations each rank writes its blocks of the
    two live Lanczos vectors (``storage.save_vector``: ``<prefix>.v.rXofP.lsb``, ``<prefix>.vprev...``) and rank 0 the
    recurrence coefficients (``<prefix>.json``); ``resume=True`` continues such a run from the last complete checkpoint
    with bit-identical coefficients (a 150-iteration kagome-42 run is 13 minutes on 8 GPUs: longer than one job slot)."""
    import time
    import torch
    sh = _wrap(operator)
    t_start = time.perf_counter()

    def save_checkpoint(done, v, v_prev, coeffs):
        import json
        from pathlib import Path
        from . import storage
        L = sh.layout
        sh.sync()
        for name, vec in (("v", v), ("vprev", v_prev)):
            tmp = storage.save_vector(f"{checkpoint}.{name}.tmp.lsb", vec, L.row_begin, L.dim, L.rank, L.world)
            tmp.replace(storage.rank_path(f"{checkpoint}.{name}.lsb", L.rank, L.world))
        if L.rank == 0:   # written last: a checkpoint counts only once its header is there
            host = coeffs[:, :done].cpu().numpy()
            meta = {"iterations": int(done), "seed": int(seed), "dim": int(L.dim), "world": int(L.world),
                    "alphas": [float.hex(float(a)) for a in host[0]], "betas": [float.hex(float(b)) for b in host[1]]}
            tmp = Path(f"{checkpoint}.json.tmp")
            tmp.write_text(json.dumps(meta))
            tmp.replace(f"{checkpoint}.json")

    def load_checkpoint():
        import json
        from pathlib import Path
        from . import storage
        L = sh.layout
        meta = json.loads(Path(f"{checkpoint}.json").read_text())
        if (meta["dim"], meta["world"], meta["seed"]) != (L.dim, L.world, seed):
            raise ValueError("the checkpoint belongs to a different run (dim / ranks / seed)")
        device = getattr(sh, "device", "cuda")
        vecs = []
        for name in ("v", "vprev"):
            a, header = storage.load_vector(f"{checkpoint}.{name}.lsb", L.rank, L.world)
            if (header["row_begin"], header["row_end"]) != (L.row_begin, L.row_end):
                raise ValueError("the checkpoint was written with a different row layout")
            vecs.append(torch.from_numpy(np.ascontiguousarray(a)).to(device))
        return meta["iterations"], vecs[0], vecs[1], [float.fromhex(a) for a in meta["alphas"]], \
            [float.fromhex(b) for b in meta["betas"]]

    def run(accumulate_with=None):
        v = _start_vector(sh, seed)
        v /= torch.sqrt(sh.dot(v, v))
        v_prev = sh.empty_vector()
        w = sh.empty_vector()
        out = sh.empty_vector() if accumulate_with is not None else None
        n_steps = len(accumulate_with) if accumulate_with is not None else max_iters
        coeffs = torch.zeros(2, max(n_steps, 1), dtype=torch.float64, device=v.device)  # alphas; betas -- on the device
        weights = None if accumulate_with is None else torch.as_tensor(np.asarray(accumulate_with), device=v.device)
        energy, resid, converged, done = float("nan"), float("inf"), False, 0
        beta = None
        history = []
        first = 0
        if resume and checkpoint is not None and accumulate_with is None:
            first, v, v_prev, alphas0, betas0 = load_checkpoint()
            coeffs[0, :first] = torch.as_tensor(alphas0, dtype=torch.float64, device=coeffs.device)
            coeffs[1, :first] = torch.as_tensor(betas0, dtype=torch.float64, device=coeffs.device)
            beta = coeffs[1, first - 1:first].clone()
            done = first
        for k in range(first, n_steps):
            if out is not None:
                out.addcmul_(v, weights[k:k + 1])
            sh.matvec(v, w)
            alpha = sh.dot(v, w)
            coeffs[0, k:k + 1] = alpha
            w.addcmul_(v, alpha, value=-1.0)
            if k > 0:
                w.addcmul_(v_prev, beta, value=-1.0)
            beta = torch.sqrt(sh.dot(w, w))
            coeffs[1, k:k + 1] = beta
            done = k + 1
            if accumulate_with is None and (done % check_every == 0 or done == n_steps):
                host = coeffs[:, :done].cpu().numpy()  # the only host read: every check_every iterations
                energy, s = _lowest(host[0], host[1])
                resid = abs(host[1, -1] * s[-1])
                if progress is not None:
                    progress(done, energy, resid)
                history.append(energy)
                if resid <= tol * max(1.0, abs(energy)) or host[1, -1] < 1e-14:
                    converged = True
                    break
                if energy_tol is not None and len(history) >= 3 and all(
                        abs(history[-i] - history[-i - 1]) <= energy_tol * max(1.0, abs(energy)) for i in (1, 2)):
                    converged = True
                    break
                if time_limit_s is not None and time.perf_counter() - t_start > time_limit_s:
                    break
            v_prev, v, w = v, w, v_prev
            v /= beta   # (a vanishing beta -- invariant subspace found -- is caught at the next check)
            if checkpoint is not None and checkpoint_every > 0 and accumulate_with is None and done % checkpoint_every == 0:
                save_checkpoint(done, v, v_prev, coeffs)
        sh.sync()
        host = coeffs[:, :done].cpu().numpy()
        return host[0].copy(), host[1].copy(), energy, resid, converged, out

    alphas, betas, energy, resid, converged, _ = run()
    result = LanczosResult(energy, len(alphas), resid, converged, None, alphas, betas, matvecs=len(alphas))
    if compute_eigenvector:
        _, s = _lowest(alphas, betas)
        *_, vec = run(accumulate_with=s)
        vec /= torch.sqrt(sh.dot(vec, vec))
        result.eigenvector = vec
        result.matvecs += len(alphas)
    return result


def lanczos_thick_restart(operator, k: int = 1, basis_size: Optional[int] = None, tol: float = 1e-10,
                          max_restarts: int = 200, seed: int = 42, dtype=None) -> LanczosResult:
    """The ``k`` lowest eigenpairs (chapel/src/Diagonalize.chpl:166-177 ``numEvals``) by thick-restart Lanczos.

    A Krylov basis of at most ``basis_size`` vectors lives in HBM ([m, local rows]); every new vector is
    orthogonalised against the whole basis (classical Gram-Schmidt, twice: two fused ``V w`` products and one
    all-reduce of m numbers each); when the basis is full the k lowest Ritz vectors are kept (one GEMM) and the
    iteration continues from the residual direction.  Works for real symmetric and complex Hermitian operators
    (``dtype=torch.complex128``).  ``tol`` bounds every residual |beta s_i[m]| relative to max(1, |theta_i|)."""
    import torch
    from . import _lib
    sh = _wrap(operator)
    dtype = dtype or torch.float64
    m = basis_size or max(2 * k + 16, 24)
    n = sh.layout.rows
    device = getattr(sh, "device", "cuda")
    V = torch.zeros(m, n, dtype=dtype, device=device)
    T = np.zeros((m, m), dtype=np.complex128 if dtype == torch.complex128 else np.float64)
    w = sh.empty_vector(dtype)

    def reduce_(t):
        if sh.layout.world > 1 and hasattr(sh, "allreduce_"):
            return sh.allreduce_(t)   # a stand-in that brings its own collective (gloo tests)
        if sh.layout.world > 1:
            view = torch.view_as_real(t) if t.dtype == torch.complex128 else t
            view = view.contiguous()
            _lib.lib.ls_b200_comm_allreduce_f64(view.data_ptr(), view.numel())
            _lib.check_error()
            if t.dtype == torch.complex128:
                t.copy_(torch.view_as_complex(view))
            else:
                t.copy_(view)
        return t

    v0 = _start_vector(sh, seed, dtype)
    V[0] = v0 / torch.sqrt(sh.dot(v0, v0).real)
    have = 1          # basis vectors present
    locked = 0        # Ritz vectors carried over from the last restart (rows 0 .. locked-1 of T are diagonal + border)
    matvecs = 0
    theta = np.zeros(k)
    resid = np.full(k, np.inf)
    converged = False
    for restart in range(max_restarts + 1):
        j = have - 1
        beta_last = 0.0
        while True:
            sh.matvec(V[j], w)
            matvecs += 1
            # coefficients against the whole basis, twice (CGS2); the first pass also yields column j of T
            h = reduce_(torch.mv(V[:have].conj(), w))
            w -= torch.mv(V[:have].t(), h)
            h2 = reduce_(torch.mv(V[:have].conj(), w))
            w -= torch.mv(V[:have].t(), h2)
            col = (h + h2).cpu().numpy()
            T[:have, j] = col
            T[j, :have] = np.conj(col)
            T[j, j] = col[j].real
            beta = float(torch.sqrt(sh.dot(w, w).real).item())
            beta_last = beta
            if have == m or beta < 1e-13:
                break
            V[have] = w / beta
            T[have, j] = T[j, have] = beta
            have += 1
            j += 1
        evals, S = np.linalg.eigh(T[:have, :have])
        kk = min(k, have)
        theta = evals[:kk]
        resid = np.abs(beta_last * S[have - 1, :kk])
        if np.all(resid <= tol * np.maximum(1.0, np.abs(theta))) or beta_last < 1e-13:
            converged = True
        if converged or restart == max_restarts:
            St = torch.as_tensor(S[:, :kk].T.copy(), device=device).to(dtype)
            vecs = St @ V[:have]
            sh.sync()
            return LanczosResult(float(theta[0]), matvecs, float(resid.max()), converged, vecs[0], None, None,
                                 energies=np.array(theta), residuals=np.array(resid), eigenvectors=vecs, matvecs=matvecs)
        # thick restart: keep `keep` Ritz vectors, continue from the residual direction w / beta
        keep = min(have - 1, max(kk + 4, (kk + have) // 2 if have > 2 * kk else kk))
        keep = max(1, min(keep, m - 2))
        St = torch.as_tensor(S[:, :keep].T.copy(), device=device).to(dtype)
        V[:keep] = St @ V[:have]
        V[keep] = w / beta_last
        T[:] = 0
        for i in range(keep):
            T[i, i] = evals[i]
            T[keep, i] = beta_last * S[have - 1, i]
            T[i, keep] = np.conj(T[keep, i])
        locked = keep
        have = keep + 1
    raise AssertionError("unreachable")


def lanczos_block_thick_restart(operator, k: int = 1, block_size: int = 4, basis_size: Optional[int] = None,
                                tol: float = 1e-10, max_restarts: int = 200, seed: int = 42, dtype=None) -> LanczosResult:
    """The ``k`` lowest eigenpairs by thick-restart BLOCK Lanczos: every step multiplies ``block_size`` vectors at once.

    The Krylov basis ([m, local rows], in HBM) grows by one block per step: W = H V_block through the operator's
    ``matvec_block`` (one pass over the matrix elements for all b vectors; falls back to b single products for complex
    vectors and sharded operators), coefficients against the whole basis twice (CGS2: two GEMMs + one all-reduce each),
    then the residual block is orthonormalised through its b x b Gram matrix (eigen-decomposition, twice; directions
    whose norm fell below 1e-7 of the largest are dropped, so converged / linearly dependent directions shrink the
    block instead of polluting the basis).  When the basis is full the lowest Ritz vectors are kept together with the
    residual block and the iteration continues (Wu-Simon restart, block form).  With ``block_size`` = 1 this is
    :func:`lanczos_thick_restart`.  ``tol`` bounds every residual norm relative to max(1, |theta_i|)."""
    import torch
    from . import _lib
    sh = _wrap(operator)
    dtype = dtype or torch.float64
    cplx = dtype == torch.complex128
    device = getattr(sh, "device", "cuda")
    n = sh.layout.rows
    dim = sh.layout.dim
    b = max(1, min(int(block_size), dim))
    k = max(1, min(int(k), dim))
    m = basis_size or max(2 * k + 4 * b, 24)
    m = min(max(m, k + 2 * b), dim)
    V = torch.zeros(m, n, dtype=dtype, device=device)
    T = np.zeros((m, m), dtype=np.complex128 if cplx else np.float64)
    W_full = torch.zeros(b, n, dtype=dtype, device=device)
    use_block = (not cplx) and hasattr(sh, "matvec_block")

    def reduce_(t):
        if sh.layout.world > 1 and hasattr(sh, "allreduce_"):
            return sh.allreduce_(t)   # a stand-in that brings its own collective (gloo tests)
        if sh.layout.world > 1:
            view = (torch.view_as_real(t) if cplx else t).contiguous()
            _lib.lib.ls_b200_comm_allreduce_f64(view.data_ptr(), view.numel())
            _lib.check_error()
            t = torch.view_as_complex(view) if cplx else view
        return t

    def gram(A, B):
        """<a_i, b_j> over all ranks for the rows of A [p, n] and B [q, n] -> [p, q] on the device."""
        return reduce_(A.conj() @ B.t())

    def orthonormalise(W):
        """Rows of W -> (Q [c, n] orthonormal, B [c, rows of W]) with W = B^T-combination of Q: w_j = sum_k B[k, j] q_k."""
        c0 = W.shape[0]
        B_total = np.eye(c0, dtype=T.dtype)
        Q = W
        for _pass in range(2):
            G = gram(Q, Q).cpu().numpy()
            G = (G + G.conj().T) / 2
            lam, U = np.linalg.eigh(G)
            top = max(float(lam[-1]), 0.0)
            keep = lam > max(top * 1e-14, 1e-280)      # norms below 1e-7 of the largest: dependent / converged directions
            lam, U = lam[keep], U[:, keep]
            if lam.size == 0:
                return Q[:0], np.zeros((0, c0), dtype=T.dtype)
            # q_k = lam_k^-1/2 sum_j U[j, k] q_j  (rows: Q' = lam^-1/2 U^T Q);  q_j = sum_k (lam^1/2 U^H)[k, j] q'_k
            M = (U / np.sqrt(lam)).T
            Q = torch.as_tensor(np.ascontiguousarray(M), device=device).to(dtype) @ Q
            B_total = (np.sqrt(lam)[:, None] * U.conj().T) @ B_total
        return Q, B_total

    # start block: hashed vectors, independent of the number of ranks
    start = torch.stack([_start_vector(sh, seed + 7919 * i, dtype) for i in range(b)])
    Q0, _ = orthonormalise(start)
    del start
    cur = Q0.shape[0]
    V[:cur] = Q0
    have = cur               # basis vectors present
    j0 = 0                   # the block still to be multiplied is V[j0:have]
    matvecs = 0
    theta = np.zeros(k)
    resid = np.full(k, np.inf)
    converged = False
    for restart in range(max_restarts + 1):
        pending_Q, pending_B = None, None
        while True:
            c = have - j0
            W = W_full[:c]
            if use_block and c > 1:
                sh.matvec_block(V[j0:have], W)
            else:
                for i in range(c):
                    sh.matvec(V[j0 + i], W[i])
            matvecs += c
            h = gram(V[:have], W)                      # [have, c]: column block j0:have of T
            W -= h.t() @ V[:have]
            h2 = gram(V[:have], W)
            W -= h2.t() @ V[:have]
            col = (h + h2).cpu().numpy()
            T[:have, j0:have] = col
            T[j0:have, :have] = col.conj().T
            blk = T[j0:have, j0:have]
            T[j0:have, j0:have] = (blk + blk.conj().T) / 2
            scale = max(1.0, float(np.abs(col).max()))
            Q, B = orthonormalise(W)
            # directions that carry nothing any more (invariant subspace reached): drop them
            live = np.linalg.norm(B, axis=1) > 1e-13 * scale if B.shape[0] else np.zeros(0, dtype=bool)
            if B.shape[0] and not live.all():
                idx = torch.as_tensor(np.nonzero(live)[0], device=device)
                Q, B = Q[idx], B[live]
            cnew = Q.shape[0]
            pending_Q, pending_B = Q, B
            if cnew == 0 or have + cnew > m:
                break
            V[have:have + cnew] = Q
            T[have:have + cnew, j0:have] = B
            T[j0:have, have:have + cnew] = B.conj().T
            j0, have = have, have + cnew
        evals, S = np.linalg.eigh(T[:have, :have])
        kk = min(k, have)
        theta = evals[:kk]
        last = S[j0:have, :kk]                        # the rows of the Ritz vectors in the last multiplied block
        resid = np.linalg.norm(pending_B @ last, axis=0) if pending_B.shape[0] else np.zeros(kk)
        if np.all(resid <= tol * np.maximum(1.0, np.abs(theta))) or pending_B.shape[0] == 0:
            converged = True
        if converged or restart == max_restarts:
            St = torch.as_tensor(S[:, :kk].T.copy(), device=device).to(dtype)
            vecs = St @ V[:have]
            sh.sync()
            return LanczosResult(float(theta[0]), matvecs, float(resid.max()), converged, vecs[0], None, None,
                                 energies=np.array(theta), residuals=np.array(resid), eigenvectors=vecs, matvecs=matvecs)
        # thick restart: the lowest Ritz vectors + the residual block
        cnew = pending_Q.shape[0]
        keep = min(have - 1, max(kk + 4, (kk + have) // 2 if have > 2 * kk else kk))
        keep = max(1, min(keep, m - cnew - b))
        St = torch.as_tensor(S[:, :keep].T.copy(), device=device).to(dtype)
        V[:keep] = St @ V[:have]
        V[keep:keep + cnew] = pending_Q
        border = pending_B @ S[j0:have, :keep]         # [cnew, keep]
        T[:] = 0
        T[np.arange(keep), np.arange(keep)] = evals[:keep]
        T[keep:keep + cnew, :keep] = border
        T[:keep, keep:keep + cnew] = border.conj().T
        j0, have = keep, keep + cnew
    raise AssertionError("unreachable")
