"""Oracle Krylov solvers with PETSc KSP semantics.

Test infrastructure only (see ``oracle/__init__.py``).

Restates what ``LinearSolver.solve`` (``cashocs/_utils/linalg.py:480-546``) asks PETSc to do
[third party, recalled -- SURVEY.md 9.4]: zero initial guess, left preconditioning,
*preconditioned* residual norm for CG / MINRES, ``KSPConvergedDefault``
(``rnorm <= max(rtol*rnorm0, atol)`` -> 2 (rtol) / 3 (atol); ``rnorm >= dtol*rnorm0`` -> -4;
``its >= max_it`` -> -3; NaN -> -9; CG: ``p.Ap <= 0`` -> -10, ``r.z < 0`` -> -8).
Defaults rtol=1e-5, atol=1e-50, dtol=1e5, max_it=1e4.  Iteration counts are
"parity unpinned" (no reference test stores them).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

CONVERGED_RTOL = 2
CONVERGED_ATOL = 3
CONVERGED_ITS = 4
DIVERGED_ITS = -3
DIVERGED_DTOL = -4
DIVERGED_BREAKDOWN = -5
DIVERGED_INDEFINITE_PC = -8
DIVERGED_NANORINF = -9
DIVERGED_INDEFINITE_MAT = -10


def _default_test(rnorm, it, state, rtol, atol, dtol):
    if it == 0:
        state["ttol"] = max(rtol * rnorm, atol)
        state["rnorm0"] = rnorm
    if rnorm != rnorm or np.isinf(rnorm):
        return DIVERGED_NANORINF
    if rnorm <= state["ttol"]:
        return CONVERGED_ATOL if rnorm < atol else CONVERGED_RTOL
    if rnorm >= dtol * state["rnorm0"]:
        return DIVERGED_DTOL
    return 0


def jacobi(A: sp.csr_matrix):
    dinv = 1.0 / A.diagonal()
    return lambda r: dinv * r


def chebyshev_jacobi(A: sp.csr_matrix, degree: int, lmax: float | None = None, ratio: float = 30.0):
    """Fixed Chebyshev polynomial in D^-1 A used as a (linear, SPD) preconditioner.

    ``z = p_k(D^-1 A) D^-1 r`` with the k-step Chebyshev iteration for eigenvalues in
    ``[lmax/ratio, lmax]``; ``lmax`` defaults to the Gershgorin bound
    ``max_i sum_j |a_ij| / a_ii`` (deterministic, no power iteration).
    """
    d = A.diagonal()
    dinv = 1.0 / d
    if lmax is None:
        lmax = float(np.max(np.asarray(abs(A).sum(axis=1)).reshape(-1) * dinv))
    lmin = lmax / ratio
    theta = 0.5 * (lmax + lmin)
    delta = 0.5 * (lmax - lmin)

    def apply(r):
        # standard Chebyshev semi-iteration for A z = r, z0 = 0 (Saad, Alg. 12.1)
        sigma = theta / delta
        rho = 1.0 / sigma
        res = r.copy()
        dvec = dinv * res / theta
        z = dvec.copy()
        for _ in range(degree - 1):
            res = r - A @ z
            rho_new = 1.0 / (2.0 * sigma - rho)
            dvec = rho_new * rho * dvec + (2.0 * rho_new / delta) * (dinv * res)
            z = z + dvec
            rho = rho_new
        return z

    return apply


def cg(A, b, pc=None, rtol=1e-5, atol=1e-50, dtol=1e5, max_it=10000, history=None):
    """PETSc ``KSPCG`` (``-ksp_norm_type preconditioned``) -> (x, reason, its, rnorm)."""
    n = b.shape[0]
    apply_pc = pc if pc is not None else (lambda r: r.copy())
    x = np.zeros(n)
    r = b.copy()
    z = apply_pc(r)
    beta = float(z @ r)
    dp = float(np.sqrt(z @ z))
    state: dict = {}
    if history is not None:
        history.append(dp)
    reason = _default_test(dp, 0, state, rtol, atol, dtol)
    if reason:
        return x, reason, 0, dp
    p = np.zeros(n)
    betaold = 1.0
    i = 0
    while True:
        if beta == 0.0:
            return x, CONVERGED_ATOL, i + 1, dp
        if beta < 0.0:
            return x, DIVERGED_INDEFINITE_PC, i + 1, dp
        if i == 0:
            p = z.copy()
        else:
            p = z + (beta / betaold) * p
        betaold = beta
        w = A @ p
        dpi = float(p @ w)
        if dpi <= 0.0 or dpi != dpi:
            return x, (DIVERGED_NANORINF if dpi != dpi else DIVERGED_INDEFINITE_MAT), i + 1, dp
        a = beta / dpi
        x = x + a * p
        r = r - a * w
        z = apply_pc(r)
        beta = float(z @ r)
        dp = float(np.sqrt(z @ z))
        if history is not None:
            history.append(dp)
        i += 1
        reason = _default_test(dp, i, state, rtol, atol, dtol)
        if reason:
            return x, reason, i, dp
        if i >= max_it:
            return x, DIVERGED_ITS, i, dp


def minres(A, b, pc=None, rtol=1e-5, atol=1e-50, dtol=1e5, max_it=10000, history=None):
    """Preconditioned MINRES (Paige & Saunders 1975), zero initial guess -> (x, reason, its, rnorm).

    The monitored quantity is the recurrence estimate of the preconditioned residual norm
    ``||r||_{M^-1}`` (``phibar``), as in PETSc's classical ``KSPMINRES`` [third party, recalled].
    """
    n = b.shape[0]
    apply_pc = pc if pc is not None else (lambda r: r.copy())
    x = np.zeros(n)
    r1 = b.copy()
    y = apply_pc(r1)
    beta1 = float(r1 @ y)
    if beta1 < 0.0:
        return x, DIVERGED_INDEFINITE_PC, 0, float("nan")
    beta1 = np.sqrt(beta1)
    state: dict = {}
    if history is not None:
        history.append(beta1)
    reason = _default_test(beta1, 0, state, rtol, atol, dtol)
    if reason:
        return x, reason, 0, beta1
    oldb, beta = 0.0, beta1
    dbar, epsln = 0.0, 0.0
    phibar = beta1
    cs, sn = -1.0, 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    r2 = r1.copy()
    i = 0
    while True:
        s = 1.0 / beta
        v = s * y
        y = A @ v
        if i >= 1:
            y = y - (beta / oldb) * r1
        alfa = float(v @ y)
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        y = apply_pc(r2)
        oldb = beta
        beta = float(r2 @ y)
        if beta < 0.0:
            return x, DIVERGED_INDEFINITE_PC, i + 1, phibar
        beta = np.sqrt(beta)
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        gamma = max(np.sqrt(gbar * gbar + beta * beta), np.finfo(float).tiny)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) / gamma
        x = x + phi * w
        i += 1
        rn = abs(phibar)
        if history is not None:
            history.append(rn)
        reason = _default_test(rn, i, state, rtol, atol, dtol)
        if reason:
            return x, reason, i, rn
        if i >= max_it:
            return x, DIVERGED_ITS, i, rn


def direct(A, b):
    """``preonly`` + ``lu`` (MUMPS in the reference, ``cashocs/_utils/linalg.py:54-59``)."""
    return spla.spsolve(A.tocsc(), b), CONVERGED_ITS, 1, 0.0


def solve(A, b, ksp_options=None, rtol=None, atol=None, history=None, block_size=1, coords=None):
    """Dispatch on a cashocs ``ksp_options`` dict like ``LinearSolver.solve``; b None -> x = 0.
    ``block_size`` / ``coords`` describe a vector problem to the multigrid preconditioner
    (``pc_type`` hypre / gamg -> ``oracle/amg.py``)."""
    if b is None:
        return np.zeros(A.shape[0]), CONVERGED_ATOL, 0, 0.0
    opts = dict(ksp_options) if ksp_options is not None else {"ksp_type": "preonly", "pc_type": "lu"}
    ksp = str(opts.get("ksp_type", "gmres"))
    pct = str(opts.get("pc_type", "none"))
    kw = {
        "rtol": float(opts.get("ksp_rtol", 1e-5)) if rtol is None else rtol,
        "atol": float(opts.get("ksp_atol", 1e-50)) if atol is None else atol,
        "dtol": float(opts.get("ksp_divtol", 1e5)),
        "max_it": int(opts.get("ksp_max_it", 10000)),
    }
    if ksp == "preonly":
        if pct == "lu":
            return direct(A, b)
        if pct == "jacobi":
            return b / A.diagonal(), CONVERGED_ITS, 1, 0.0
        raise ValueError(f"oracle: unsupported preonly pc_type {pct}")
    if pct == "jacobi":
        pc = jacobi(A)
    elif pct == "none":
        pc = None
    elif pct == "chebyshev":
        pc = chebyshev_jacobi(A, int(opts.get("pc_chebyshev_degree", 3)))
    elif pct in ("hypre", "gamg", "amg"):
        from oracle import amg

        pc = amg.SmoothedAggregation(A, bs=block_size, coords=coords)
    else:
        raise ValueError(f"oracle: unsupported pc_type {pct}")
    if ksp == "cg":
        return cg(A, b, pc, history=history, **kw)
    if ksp == "minres":
        return minres(A, b, pc, history=history, **kw)
    raise ValueError(f"oracle: unsupported ksp_type {ksp}")
