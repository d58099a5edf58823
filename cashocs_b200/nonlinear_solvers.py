"""``cashocs.linear_solve`` / ``cashocs.newton_solve`` (``cashocs/nonlinear_solvers/linear_solver.py:37-110``,
``newton_solver.py:365-470``) for the form family of the device layer: assemble (``cb_assemble_scalar``,
``cb_assemble_cubic``) and solve (``cb_ksp_solve``) one stand-alone PDE, the same primitives the state problem uses.

``form`` is a :class:`PoissonForm` whose unknown is ``u`` (``kappa grad u . grad v + reaction u v + cubic u^3 v -
source v``); ``bcs`` a list of :class:`DirichletBC`.  Options that select machinery outside the hot path (custom
derivative / shift forms, preconditioner forms, damped or inexact Newton variants other than the full step) raise.
"""

from __future__ import annotations

import copy

from cashocs_b200 import _exceptions
from cashocs_b200._forms import forms
from cashocs_b200._pde_problems import state_problem
from cashocs_b200._utils import linalg
from cashocs_b200.io import config as config_module


def _enlist(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _problem(form, u, bcs, ksp_options, linear_solver, is_linear: bool, rtol: float, atol: float, max_iter: int):
    if not isinstance(form, forms.PoissonForm):
        raise _exceptions.InputError("cashocs_b200.nonlinear_solvers", "form", "a PoissonForm of the supported family")
    cfg = config_module.Config()
    cfg.set("StateSystem", "is_linear", str(bool(is_linear)))
    cfg.set("StateSystem", "newton_rtol", repr(float(rtol)))
    cfg.set("StateSystem", "newton_atol", repr(float(atol)))
    cfg.set("StateSystem", "newton_iter", str(int(max_iter)))
    options = copy.deepcopy(linalg.direct_ksp_options) if ksp_options is None else ksp_options
    u.mesh.build_pattern()
    return state_problem.StateProblem(u.mesh, [form], [_enlist(bcs)], [u], [options], cfg, linear_solver)


def linear_solve(linear_form, u, bcs=None, ksp_options=None, rtol: float | None = None, atol: float | None = None,
                 A_tensor=None, b_tensor=None, preconditioner_form=None, linear_solver=None):
    """Solves a linear problem (``linear_solver.py:37-110``); the solution is written into ``u`` and returned."""
    if preconditioner_form is not None or A_tensor is not None or b_tensor is not None:
        raise _exceptions.InputError("cashocs_b200.linear_solve", "preconditioner_form",
                                     "separate preconditioner forms / user tensors are not supported")
    if getattr(linear_form, "is_nonlinear", False):
        raise _exceptions.InputError("cashocs_b200.linear_solve", "linear_form", "the form is nonlinear: use newton_solve")
    prob = _problem(linear_form, u, bcs, ksp_options, linear_solver, True, 1e-10, 1e-10, 50)
    A, b = prob.assemble(0)
    res = prob.linear_solver.solve(u.vector, A=A, b=b, ksp_options=prob.ksp_options[0], rtol=rtol, atol=atol)
    if u.mesh.dist is not None:
        u.mesh.dist.halo_exchange(u.vector, 1)
    u.iterations = res.its
    return u


def newton_solve(nonlinear_form, u, bcs=None, derivative=None, shift=None, rtol: float = 1e-10, atol: float = 1e-10,
                 max_iter: int = 50, convergence_type: str = "combined", norm_type: str = "l2", damped: bool = False,
                 inexact: bool = False, verbose: bool = True, ksp_options=None, A_tensor=None, b_tensor=None,
                 is_linear: bool = False, preconditioner_form=None, linear_solver=None):
    """Newton's method for ``F(u) = 0`` (``newton_solver.py:285-362,365-470``) with full steps and the combined
    convergence test ``||F(u)|| <= atol + rtol ||F(u_0)||`` in the l2 norm; ``u`` is the initial guess and the result."""
    for name, value, ok in (("derivative", derivative, None), ("shift", shift, None), ("A_tensor", A_tensor, None),
                            ("b_tensor", b_tensor, None), ("preconditioner_form", preconditioner_form, None)):
        if value is not ok:
            raise _exceptions.InputError("cashocs_b200.newton_solve", name, "not supported on the device path")
    if convergence_type != "combined" or norm_type != "l2" or damped or inexact:
        raise _exceptions.InputError("cashocs_b200.newton_solve", "convergence_type",
                                     "the device path runs undamped, exact Newton steps with the combined l2 test")
    prob = _problem(nonlinear_form, u, bcs, ksp_options, linear_solver, False, rtol, atol, max_iter)
    prob.solve()
    u.iterations = prob.newton_iterations[0]
    return u
