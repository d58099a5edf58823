"""Newton's method for semilinear state equations (TEST INFRASTRUCTURE ONLY, see ``oracle/__init__.py``).

Restates ``cashocs.newton_solve`` (``cashocs/nonlinear_solvers/newton_solver.py:285-362``) for residuals of the form
``F(u)[v] = int grad u . grad v + c u^3 v - f v dx`` -- the nonlinearity of the reference's own tests
(``tests/test_nonlinear_solvers.py:28-61``: c = 1e2, f = 1; ``tests/test_optimal_control.py:364-376``: c = 1, f = control):

* the iterate carries the Dirichlet values from the start, the Newton increment has homogeneous conditions;
* ``res_0 = ||F(u_0)||`` (l2 norm of the assembled residual with the boundary rows zeroed), convergence
  ``res <= atol + rtol * res_0`` ("combined"), at most ``max_iter`` iterations, ``NotConvergedError`` otherwise;
* undamped: ``u <- u - J(u)^-1 F(u)`` with ``J(u) = K + 3 c M[u^2]``.

The cubic load ``int u^3 phi_a`` and the weighted mass ``int u^2 phi_a phi_b`` are integrated with a degree-4 rule,
exact for P1 ``u``.  "parity unpinned" for iteration counts; the reference's acceptance criteria (solution equals
FEniCS' own Newton solve, Taylor rate > 1.9 for the control gradient) are reproduced in ``tests/test_oracle.py``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse.linalg as spla

from oracle import fem


def cubic_load_elements(vol, cells, u, d, c):
    """b_a = c |K| sum_q w_q u(x_q)^3 phi_a(x_q)."""
    lam, w = fem.quadrature(d, 4)
    uq = u[cells] @ lam.T                                   # [Nc, nq]
    return c * vol[:, None] * np.einsum("cq,q,qa->ca", uq**3, w, lam, optimize=True)


def weighted_mass_elements(vol, cells, u, d, c):
    """M_ab = 3 c |K| sum_q w_q u(x_q)^2 phi_a phi_b  (the derivative of the cubic term)."""
    lam, w = fem.quadrature(d, 4)
    uq = u[cells] @ lam.T
    return 3.0 * c * vol[:, None, None] * np.einsum("cq,q,qa,qb->cab", uq**2, w, lam, lam, optimize=True)


def newton_solve(coords, cells, load_elements, c, bc_dofs, bc_value=0.0, u0=None, rtol=1e-11, atol=1e-13, max_iter=50):
    """Solve ``K u + c int u^3 phi = load`` with ``u = bc_value`` on ``bc_dofs``; returns (u, iterations)."""
    d = coords.shape[1]
    nv = coords.shape[0]
    _, vol, G = fem.cell_geometry(coords, cells)
    Ke = fem.stiffness_elements(vol, G)
    u = np.zeros(nv) if u0 is None else np.array(u0, dtype=float, copy=True)
    u[bc_dofs] = bc_value

    def residual(u):
        re = np.einsum("cab,cb->ca", Ke, u[cells]) + cubic_load_elements(vol, cells, u, d, c) - load_elements
        r = fem.assemble_vector(re, cells, nv)
        r[bc_dofs] = 0.0
        return r

    r = residual(u)
    res0 = float(np.linalg.norm(r))
    if res0 == 0.0:
        return u, 0
    tol = atol + rtol * res0
    res, it = res0, 0
    while res > tol and it < max_iter:
        it += 1
        Je = Ke + weighted_mass_elements(vol, cells, u, d, c)
        J, rhs = fem.assemble_system(Je, None, cells, nv, bc_dofs, 0.0)
        rhs = -r
        rhs[bc_dofs] = 0.0
        u = u + spla.spsolve(J.tocsc(), rhs)
        r = residual(u)
        res = float(np.linalg.norm(r))
    if not res <= tol:
        raise RuntimeError("NotConvergedError: Newton solver, Maximum number of iterations were exceeded.")
    return u, it
